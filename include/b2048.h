/*
 * b2048.h — C-ABI of the B200-native 2048 / Double-DQN hot path.
 *
 * This is the drop-in boundary for the reference's `board.Board2048` environment step and the
 * `dqn_lib` experience/update loop (ribal-aladeeb/reinforcement-learning-2048).  The reference has
 * no FFI of its own (it is pure Python), so every entry point below cites the
 * reference Python it replaces (paths relative to the reference root).  The Python shims
 * `board.py` / `dqn_lib.py` in this repo bind these symbols with ctypes.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - pointers are DEVICE pointers owned by the caller unless the name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - return value is an `int`: 0 = ok, >0 = a `cudaError_t`, <0 = a B2048_E* code below;
 *   - device-pointer calls never allocate, never synchronise and never throw after
 *     `b2048_init(device)`; they are safe to capture in a CUDA graph;
 *   - device-pointer calls are re-entrant: every buffer a launch touches is passed by the caller (the
 *     only library-owned device state is the read-only row table), so any number of calls may be in
 *     flight on different streams / host threads of one device.  `*_host` calls serialise per device;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Packed board format ("u64 board"): 16 tile exponents of 4 bits each; cell (row r, column c)
 * lives in nibble 4*r + c, nibble 0 = least-significant.  Exponent 0 = empty, e>0 = tile 2^e
 * (this is exactly the reference's `log_scale()` value, src/board.py:224-231).  Row-major nibble
 * order makes "k-th empty cell" agree with the reference's `np.where(state == 0)` order
 * (src/board.py:46-48).
 *
 * Action encoding everywhere: 0 = up, 1 = down, 2 = left, 3 = right (src/board.py:129,191).
 */
#ifndef B2048_H_
#define B2048_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: spawn stream v2 (one 16-bit Philox lane per board, eight boards per call; see b2048_step);
 *    boards / next of b2048_step and next4 of b2048_step_all4 take the fast path when 32-byte aligned. */
#define B2048_ABI_VERSION 2

/* ---- error codes (negative; positive values are cudaError_t) ------------------------------ */
#define B2048_OK 0
#define B2048_ENOTINIT (-1)  /* b2048_init(device) has not been called for the current device */
#define B2048_EINVAL   (-2)  /* bad argument (NULL pointer, negative size, bad device, ...)    */
#define B2048_ENODEV   (-3)  /* no usable CUDA device                                         */

/* ---- flags byte written by the step / legal-mask kernels ---------------------------------- */
#define B2048_FLAG_UP       0x01u /* bits 0-3: legal-move mask of the INPUT board,            */
#define B2048_FLAG_DOWN     0x02u /*   order [up, down, left, right]                          */
#define B2048_FLAG_LEFT     0x04u /*   (src/board.py:128-135)                                 */
#define B2048_FLAG_RIGHT    0x08u
#define B2048_FLAG_DONE     0x10u /* INPUT board has no legal move (src/dqn_lib.py:17-18)     */
#define B2048_FLAG_CHANGED  0x20u /* the move changed the board => a tile was spawned         */
#define B2048_FLAG_OVERFLOW 0x40u /* a 32768+32768 merge happened: 2^16 does not fit 4 bits;  */
                                  /*   next/reward for that board are unspecified             */
#define B2048_FLAG_BADSPAWN 0x80u /* spawn_override named a non-empty cell (nothing spawned)  */

/* spawn_override byte: low nibble = cell index 4*r+c, high nibble = exponent (1 => "2",
 * 2 => "4"); 0xFF = no override for this board (use the Philox stream); 0xFE = spawn nothing
 * (slide-only result, e.g. Board2048._apply_action_to_vector). */
#define B2048_SPAWN_NONE 0xFFu
#define B2048_SPAWN_SKIP 0xFEu

/* p4_threshold: a spawned tile is a "4" iff frac < p4_threshold (32-bit unsigned), where frac is the
 * board's 16-bit uniform fraction in the upper half of a 32-bit word (see b2048_step: Spawn).
 * 0x1999999A = 10 % (north_star), 0x80000000 = 50 % (the reference, src/board.py:12,49). */
#define B2048_P4_TEN_PERCENT   0x1999999Au
#define B2048_P4_FIFTY_PERCENT 0x80000000u

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Build the 65536-entry row table (canonical left slide+merge, = src/board.py:92-126 for every
 * row of 4-bit exponents) and upload it to `device`.  Idempotent, thread-safe.  Must precede every
 * other call on `device`. */
int b2048_init(int device);
int b2048_shutdown(int device);
int b2048_abi_version(void);
const char* b2048_error_string(int code);

/* Copy the 65536 x u32 row table to host memory (tests compare it with the oracle).
 * entry = result_row16 | (merge_reward/4) << 16 | overflow << 31. */
int b2048_copy_row_lut_host(uint32_t* out65536);

/* ---- K1: environment step ------------------------------------------------------------------ */

/* One action per board.  Replaces Board2048.peek_action (src/board.py:185-202) -> up/down/left/
 * right (:147-183) -> _apply_action_to_vector (:92-126) -> _populate_empty_cell (:41-51),
 * dqn_lib.reward_func_merge_score (src/dqn_lib.py:87-88) and the legal-mask / done test of
 * epsilon_greedy_policy (src/dqn_lib.py:17-18) for n boards at once.
 *   next[i]   = slide/merge of boards[i] by actions[i], plus one spawned tile iff changed
 *   reward[i] = sum of merged tile values of this move
 *   flags[i]  = B2048_FLAG_* (legal mask and done are properties of the INPUT board)
 * Spawn (stream v2): board g = index_base + i owns one 16-bit lane d of the Philox4x32-7 call with key
 * `seed` and counter (g >> 3, `step`): lane g & 7 = half (g & 1) of output word ((g & 7) >> 1), low
 * half first.  With n = number of empty cells of the slid board: cell = k-th empty cell (row-major),
 * k = floor(d * n / 65536); value "4" iff (((d * n) mod 65536) << 16) < p4_threshold.  (One 32x32->64
 * product (d << 16) * n yields k in its high and the fraction in its low word.  For odd n the
 * fraction is exactly uniform, for even n uniform over multiples of n's power of two; a cell's
 * probability differs from 1/n by at most 2^-16.  Seven rounds: the smallest Crush-resistant Philox4x32
 * (Salmon et al., SC'11; Random123's philox4x32_7); every other stream of this library uses ten.)
 * The result depends only on (seed, step, g, board, action), never on n or on the sharding.
 * spawn_override (nullable, n bytes) replays a given (cell, value) instead — the parity hook.
 * actions[i] > 3 is treated as actions[i] & 3. */
int b2048_step(const uint64_t* boards, const uint8_t* actions, uint64_t* next, int32_t* reward,
               uint8_t* flags, int64_t n, uint64_t seed, uint64_t step, uint64_t index_base,
               uint32_t p4_threshold, const uint8_t* spawn_override, void* stream);

/* All four actions per board (BASELINE.json config 2; = Board2048.available_moves,
 * src/board.py:138-145).  next4[i*4+a], reward4[i*4+a]; flags[i] bits 0-3 = legal mask, bit 4 =
 * done, bit 6 = overflow in any direction.  Successor a equals b2048_step's output for action a
 * (same Philox lane).  next4 must be 32-byte aligned, reward4 16-byte aligned (B2048_EINVAL
 * otherwise); spawn_override4 (nullable) is n*4 bytes, 4-byte aligned. */
int b2048_step_all4(const uint64_t* boards, uint64_t* next4, int32_t* reward4, uint8_t* flags,
                    int64_t n, uint64_t seed, uint64_t step, uint64_t index_base,
                    uint32_t p4_threshold, const uint8_t* spawn_override4, void* stream);

/* Legal mask + done only (= available_moves_as_torch_unit_vector, src/board.py:128-135). */
int b2048_legal_mask(const uint64_t* boards, uint8_t* flags, int64_t n, void* stream);

/* Fresh boards: zeros + two spawns (= Board2048.__init__, src/board.py:10-20).  If `where_flags`
 * is non-NULL only boards with (where_flags[i] & B2048_FLAG_DONE) are reset (episode auto-reset,
 * src/dqn_lib.py:176). */
int b2048_reset(uint64_t* boards, int64_t n, uint64_t seed, uint64_t step, uint64_t index_base,
                uint32_t p4_threshold, const uint8_t* where_flags, void* stream);

/* One new tile on every board that has an empty cell (= Board2048._populate_empty_cell,
 * src/board.py:41-51), in place, with the same Philox word assignment as b2048_step.  If
 * `where_flags` is non-NULL only boards with (where_flags[i] & B2048_FLAG_CHANGED) get a tile. */
int b2048_spawn(uint64_t* boards, int64_t n, uint64_t seed, uint64_t step, uint64_t index_base,
                uint32_t p4_threshold, const uint8_t* where_flags, void* stream);

/* Per-game bookkeeping of a batched rollout, fused with the episode auto-reset (the fields the
 * reference hands to Experiment.add_episode, src/experiments.py:112-122, kept on the device).
 * For every board i, after b2048_step produced (next, reward, flags) from `prev`:
 *   ep_score[i] += reward[i]; ep_moves[i] += 1; ep_qsum[i] += max_q[i] (if max_q != NULL);
 *   if flags[i] & DONE: totals {games, sum of merge scores, sum of moves, sum of mean max-Q} and
 *   max_tile_hist[exponent of the largest tile of prev[i]] are updated, the accumulators are
 *   cleared and next[i] is replaced by a fresh board (zeros + two spawns, as b2048_reset).
 * totals: device int64[4] {games, score_sum, moves_sum, reserved}; qmean_sum: device double[1];
 * max_tile_hist: device int64[16].  All accumulators are caller-allocated and zero-initialised. */
int b2048_episode_end(uint64_t* next, const uint64_t* prev, const int32_t* reward,
                      const uint8_t* flags, const double* max_q, int64_t* ep_score, int32_t* ep_moves,
                      double* ep_qsum, int64_t* totals, double* qmean_sum, int64_t* max_tile_hist,
                      int64_t n, uint64_t seed, uint64_t step, uint64_t index_base,
                      uint32_t p4_threshold, void* stream);

/* int64 tiles [n,16] (reference `state`, row-major) <-> packed boards.  bad[i] (nullable) is set
 * to 1 if a tile is not 0 or a power of two in 2..32768. */
int b2048_pack(const int64_t* tiles, uint64_t* boards, uint8_t* bad, int64_t n, void* stream);
int b2048_unpack_tiles(const uint64_t* boards, int64_t* tiles, int64_t n, void* stream);
/* Network input layout: exponents as float64, [n,16] == [n,1,4,4] contiguous
 * (= log_scale().state_as_4d_tensor() / flattened_state_as_tensor(), src/board.py:224-237). */
int b2048_unpack_f64(const uint64_t* boards, double* out, int64_t n, void* stream);

/* Synthetic inputs of SURVEY.md §8(d): each cell empty with probability p_empty_threshold/2^32,
 * else exponent uniform in 1..max_exp; actions uniform in 0..3. */
int b2048_random_boards(uint64_t* boards, int64_t n, uint64_t seed, uint64_t index_base,
                        uint32_t p_empty_threshold, uint32_t max_exp, void* stream);
int b2048_random_actions(uint8_t* actions, int64_t n, uint64_t seed, uint64_t step,
                         uint64_t index_base, void* stream);

/* ONE board per call, host data in and out: the engine behind the drop-in `board.Board2048` (BASELINE config 1:
 * player.py's loop calls the board one at a time).  One kernel launch and one stream synchronisation per call:
 * the 16 tile values travel as a kernel argument, the result is written to mapped pinned memory.
 *   op 0 MOVE  : next[0] = move of `tiles16` by `action` (+ spawn iff spawn != 0 and the board changed), reward[0],
 *                flags = B2048_FLAG_* of b2048_step                       (src/board.py:147-202)
 *   op 1 ALL4  : next[a], reward[a] for the four actions, flags as b2048_step_all4   (src/board.py:138-145)
 *   op 2 LEGAL : flags = legal mask | done                                 (src/board.py:128-135)
 *   op 3 SPAWN : next[0] = tiles16 with one spawned tile                   (src/board.py:41-51)
 *   op 4 FRESH : next[0] = a fresh board (two spawns; tiles16 ignored)     (src/board.py:10-20)
 * Same arithmetic and Philox lanes as the batched entry points with n = 1, index_base = 0.  result->bad != 0:
 * a tile value was not 0 or a power of two in 2..32768 (nothing else is written).  Calls on one device serialise. */
typedef struct b2048_board_result {
  int64_t next[4][16];
  int32_t reward[4];
  uint32_t flags;
  uint32_t bad;
} b2048_board_result;
int b2048_board_host(int op, const int64_t* tiles16, int action, int spawn, uint64_t seed, uint64_t step,
                     uint32_t p4_threshold, b2048_board_result* result, int device);

/* Pinned host memory for b2048_step_host, placed on the NUMA node next to `device`: the calling thread is
 * bound to the GPU's local CPUs (sysfs local_cpulist) while the pages are allocated and first touched, then
 * its affinity is restored.  *numa_node_out = that node (-1 unknown), *bound_out = 1 if the binding was
 * applied (both nullable).  Falls back to a plain cudaHostAlloc when the topology cannot be read. */
int b2048_host_alloc(void** out, size_t bytes, int device, int* numa_node_out, int* bound_out);
int b2048_host_free(void* p);
/* Bind the calling host thread to the CPUs next to `device` (the thread that calls b2048_step_host). */
int b2048_bind_thread_near(int device);

/* Same as b2048_step but with HOST buffers: chunks the batch and overlaps H2D copy, kernel and
 * D2H copy on internal streams; returns after everything has landed in the host buffers.
 * Allocates its device workspace lazily on first use (per device). */
int b2048_step_host(const uint64_t* h_boards, const uint8_t* h_actions, uint64_t* h_next,
                    int32_t* h_reward, uint8_t* h_flags, int64_t n, uint64_t seed, uint64_t step,
                    uint64_t index_base, uint32_t p4_threshold, const uint8_t* h_spawn_override,
                    int device);

/* ---- K2: GPU-resident replay ring ------------------------------------------------------------ */

/* Struct-of-arrays ring of `capacity` transitions (= the deque(maxlen) of 5-tuples,
 * src/dqn_lib.py:106,172).  All arrays are caller-allocated device memory; `head_size` is a
 * device int64[4] {head, size, auto sample counter, reserved}, zero-initialised by the caller. */
typedef struct b2048_ring {
  uint64_t* s;        /* [capacity] packed state                 */
  uint64_t* s2;       /* [capacity] packed next state            */
  int32_t*  r;        /* [capacity] reward                       */
  uint8_t*  a;        /* [capacity] action                       */
  uint8_t*  d;        /* [capacity] done (0/1)                   */
  int64_t*  head_size;/* device int64[4]: next write slot, number of valid entries,            */
                      /*   automatic sample counter (see replay_sample), reserved              */
  int64_t   capacity;
} b2048_ring;

/* Append n transitions in order (oldest first); if n > capacity only the last `capacity` survive,
 * like deque(maxlen).  `done_flags` holds the step kernel's flags bytes: done = byte & 0x10
 * (B2048_FLAG_DONE); every other bit is ignored. */
int replay_append(const b2048_ring* ring, const uint64_t* s, const uint8_t* a, const int32_t* r,
                  const uint64_t* s2, const uint8_t* done_flags, int64_t n, void* stream);

/* Fused uniform sampling (with replacement) + gather + unpack into the network layout
 * (= sample_experiences + extract_samples_conv/dense, src/dqn_lib.py:33-84).
 * Logical index 0 = oldest entry, like deque indexing.  idx_override (nullable, int64[B]) replays
 * the reference's np.random.randint draw; idx_out (nullable) receives the indices used.
 * states/next_states: f64 [B,16]; actions/rewards/dones: int64 [B].
 * Sample j uses word (j & 3) of the Philox4x32-10 call with key `seed` and counter (j >> 2, ctr):
 * index = floor(word * size / 2^32).  ctr == B2048_CTR_AUTO takes the counter from the ring's
 * device-side auto counter and increments it afterwards, so a captured CUDA graph draws a fresh
 * batch on every replay. */
#define B2048_CTR_AUTO 0xFFFFFFFFFFFFFFFFull
int replay_sample(const b2048_ring* ring, int64_t B, uint64_t seed, uint64_t ctr,
                  const int64_t* idx_override, double* states, double* next_states,
                  int64_t* actions, int64_t* rewards, int64_t* dones, int64_t* idx_out,
                  void* stream);

/* ---- K3: fused Double-DQN target + summed-MSE ------------------------------------------------ */

/* = src/dqn_lib.py:125-158.  a* = argmax_j q_next_online[i,j] (first index on ties);
 * target[i] = rewards[i] + (double)((float)(1-dones[i]) * gamma_f32) * q_next_target[i,a*]
 * (use_double != 0) or ... * max_j q_next_target[i,j] (use_double == 0; q_next_online may be NULL);
 * q_sa[i] = q_cur[i, actions[i]]; loss[0] = sum_i (q_sa[i]-target[i])^2 (deterministic order);
 * grad_q_cur (nullable) [B,4] = d loss / d q_cur = 2 (q_sa - target) at column actions[i], else 0.
 * gamma is deliberately a float: the reference rounds it to float32.
 * One launch = one thread-block cluster whose partial sums meet through distributed shared memory: no
 * library-owned scratch, so concurrent calls on different streams cannot interfere. */
int ddqn_target_loss(const double* q_next_online, const double* q_next_target, const double* q_cur,
                     const int64_t* actions, const int64_t* rewards, const int64_t* dones,
                     float gamma_f32, int use_double, double* target, double* q_sa, double* loss,
                     double* grad_q_cur, int64_t B, void* stream);

/* Patch gather / scatter (im2col / col2im) for the conv Q-network's tiny convolutions written as GEMMs
 * (configs/double_dqn_conv.py:19-28: kernel_size 2, stride 1, no padding).  Activations are row
 * matrices [n*h*w, c] — what a GEMM over patches produces, and for c = 1 the same bytes as NCHW — so
 * consecutive convolutions need no layout round trip.
 *   conv_patches_f64:      x [n*h*w, c] -> cols [n*oh*ow, c*kh*kw], (c,kh,kw) order like conv.weight
 *   conv_patches_grad_f64: d cols -> d x [n*h*w, c] (each input element sums the <= kh*kw patches
 *                          that read it: a gather, deterministic, no atomics) */
int conv_patches_f64(const double* x, double* cols, int64_t n, int c, int h, int w, int kh, int kw,
                     void* stream);
int conv_patches_grad_f64(const double* dcols, double* dx, int64_t n, int c, int h, int w, int kh,
                          int kw, void* stream);

/* Fused Adam step on one flat float64 parameter buffer (= optimizer.step() of torch.optim.Adam
 * without weight decay / amsgrad, configs/double_dqn_*.py: Adam(lr=1e-2); src/dqn_lib.py:163).
 *   t = step_counter[0] + 1;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g
 *   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps);  step_counter[0] = t
 * `step_counter` is a device int64 so that a captured CUDA graph advances it on every replay. */
int ddqn_adam_step(double* params, const double* grads, double* exp_avg, double* exp_avg_sq,
                   int64_t* step_counter, int64_t n, double lr, double beta1, double beta2,
                   double eps, void* stream);

/* Gradient allreduce fused with the Adam step over NVLink peer memory (one node, one process per
 * GPU).  Every rank exposes its flat gradient buffer and a flag block to its peers (CUDA IPC);
 * one kernel per rank then (1) signals "my gradient is complete" to every peer and waits for
 * theirs, (2) sums the W peer buffers element by element in rank order — the same order on every
 * rank, so all ranks compute bit-identical sums and the replicas never drift — and applies the
 * Adam update of ddqn_adam_step to its own parameter copy, (3) signals "done reading" and waits
 * for the peers before it exits, so the next backward pass may overwrite the gradient buffers.
 *   peer_grads : device array [world] of pointers to every rank's gradient buffer (n doubles)
 *   peer_flags : device array [world] of pointers to every rank's flag block (uint64[2*world],
 *                zero-initialised); my block is peer_flags[rank]
 *   sync_state : device uint64[4], zero-initialised: {epoch, blocks-done counter, error, -}
 * Waits are bounded (30 s of wall clock); when one expires sync_state[2] is set to 1 (sticky) and the
 * kernel returns WITHOUT updating params / exp_avg / exp_avg_sq / step_counter, so a lost peer can never
 * make this replica apply a partial sum; the caller must poll sync_state[2] and treat it as fatal. */
/* Map a peer's cudaMalloc allocation into this process for access from the CURRENT device
 * (cudaIpcOpenMemHandle with lazy peer access; no context is created on the peer's device).
 * `handle64` is the 64-byte cudaIpcMemHandle_t of the allocation; *out is its base address. */
int p2p_open_ipc_handle(const unsigned char* handle64, void** out);
/* The 64-byte cudaIpcMemHandle_t of the cudaMalloc allocation that contains device pointer `ptr`. */
int p2p_get_ipc_handle(const void* ptr, unsigned char* handle64);

int p2p_allreduce_adam_f64(const double* const* peer_grads, uint64_t* const* peer_flags,
                           uint64_t* sync_state, int rank, int world, double* params,
                           double* exp_avg, double* exp_avg_sq, int64_t* step_counter, int64_t n,
                           double lr, double beta1, double beta2, double eps, void* stream);

/* ---- K8: float64 tensor-core layers of the dense Q-network (src/configs/double_dqn_dense.py:7-15) -------------- */

/* nn.Linear forward, optionally fused with ReLU: c[rows, n_out] = act(a[rows, n_in] w[n_out, n_in]^T + bias[n_out]).
 * Row-major float64, 16-byte aligned pointers, n_in even; n_out even, or n_out == 4 with relu == 0 (the Q-value
 * layer).  Replaces model(states) / model(next_states) / target_model(next_states), src/dqn_lib.py:126-147. */
int dense_linear_forward_f64(const double* a, const double* w, const double* bias, double* c, int64_t rows, int n_in,
                             int n_out, int relu, void* stream);
/* Input gradient of a Linear layer fused with the ReLU mask of the layer below:
 * dz[rows, n_in] = (g[rows, n_out] w[n_out, n_in]) * (h[rows, n_in] > 0), h = that layer's (post-ReLU) output;
 * h == NULL gives the plain product (the conv Q-network's patch-matrix gradient, masked later). */
int dense_linear_dgrad_f64(const double* g, const double* w, const double* h, double* dz, int64_t rows, int n_in,
                           int n_out, void* stream);
/* The same product with the result regrouped: input unit j = c * group + t of row i is stored at row i * group + t,
 * column c of dz [rows * group, n_in / group] (group even, n_in % group == 0, h required).  For the layer behind
 * nn.Flatten of a [channels, positions] map (src/configs/double_dqn_conv.py:24-25) with group = positions this
 * writes the gradient directly as the (board, position) x channel row matrix the convolution's backward reads. */
int dense_linear_dgrad_regroup_f64(const double* g, const double* w, const double* h, double* dz, int64_t rows,
                                   int n_in, int n_out, int group, void* stream);
/* Weight and bias gradient: dw[n_out, n_in] = g^T x, db[n_out] = column sums of g over `rows` rows (x = the layer's
 * input).  The rows are split over the SMs; per-split products go to `scratch`
 * (dense_linear_wgrad_scratch_elems(rows, n_in, n_out) doubles, 16-byte aligned) and are added in a fixed order,
 * so results are bit-reproducible.  Overwrites dw / db (no zeroing needed). */
int64_t dense_linear_wgrad_scratch_elems(int64_t rows, int n_in, int n_out);
int dense_linear_wgrad_f64(const double* g, const double* x, double* dw, double* db, double* scratch, int64_t rows,
                           int n_in, int n_out, void* stream);

/* ---- K0: batched epsilon-greedy ---------------------------------------------------------------- */

/* = epsilon_greedy_policy, src/dqn_lib.py:16-30, for n boards.  With probability eps the action is
 * uniform in 0..3 ignoring legality and max_q = 0; otherwise
 * action = argmax_j (legal_j ? (q_j - min(q)*max(q) - min(q)) : 0*(...)) with first-index ties and
 * max_q = max_j q_j.  flags = step/legal-mask flags bytes (bits 0-3 legal).
 * override (nullable, n bytes): 0xFF = draw from Philox, 0x80 = force greedy, 0..3 = force that
 * random action — the parity hook for np.random.rand()/randint (src/dqn_lib.py:20-21). */
int egreedy_select(const double* q, const uint8_t* flags, double eps, uint64_t seed, uint64_t ctr,
                   uint64_t index_base, const uint8_t* override_bytes, uint8_t* actions,
                   double* max_q, int64_t n, void* stream);

/* ---- K6: fused forward of the convolutional Q-network (no gradient) ---------------------------- */

/* = model(state) for the conv config, src/configs/double_dqn_conv.py:19-28
 * (Conv2d(1,64,2) ReLU Conv2d(64,64,2) ReLU Flatten Linear(256,64) ReLU Linear(64,4), float64), as
 * called without gradient in epsilon_greedy_policy (src/dqn_lib.py:24-25), Player.play_game
 * (src/player.py:47) and for Q(s') in train_step (src/dqn_lib.py:126-128).  One kernel, FP64
 * tensor cores.  Exactly one of `boards` (packed, n) and `states` (float64 [n,16], the layout
 * replay_sample / b2048_unpack_f64 write) is non-NULL.  scaling (boards only): 0 = exponents as
 * board.log_scale() (src/board.py:224-231), 1 = tile / largest tile as board.normalized()
 * (src/board.py:218-222).  Weights are the module's own parameter tensors, contiguous float64:
 * w1[64,1,2,2] b1[64] w2[64,64,2,2] (16-byte aligned) b2[64] w3[64,256] b3[64] w4[4,64] b4[4].
 * q: float64 [n,4].
 * Summation order differs from cuBLAS/cuDNN (agreement ~1e-13 relative). */
int qnet_conv_forward_f64(const uint64_t* boards, const double* states, int scaling, const double* w1,
                          const double* b1, const double* w2, const double* b2, const double* w3,
                          const double* b3, const double* w4, const double* b4, double* q, int64_t n,
                          void* stream);

/* The same forward for Q(s) of train_step (src/dqn_lib.py:146-150), where a backward pass follows: it
 * also writes what that pass needs — patches2 [4n,256] = the second convolution's input in im2col form
 * (row = board*4 + output position, column = channel*4 + tap; post-ReLU, exactly the operand the kernel
 * builds on the fly), act2 [n,256] = relu(conv2) in nn.Flatten order, act3 [n,64] = relu(fc1).
 * states: float64 [n,16].  The backward is layer_wgrad_small_f64 / layer_wgrad64_f64 on these plus
 * cuBLAS DGEMMs for the input gradients (b2048/qfused.py). */
int qnet_conv_forward_train_f64(const double* states, const double* w1, const double* b1, const double* w2,
                                const double* b2, const double* w3, const double* b3, const double* w4,
                                const double* b4, double* q, double* patches2, double* act2, double* act3,
                                int64_t n, void* stream);

/* The three forwards of ONE Double-DQN update as a single launch (src/dqn_lib.py:126-128 and :146):
 *   q[n,4] = online(states) + patches2 / act2 / act3 as qnet_conv_forward_train_f64 stores them,
 *   q_next_online[n,4] = online(next_states)   (NULL for plain DQN),
 *   q_next_target[n,4] = target(next_states).
 * online / target: the eight parameter pointers {w1, b1, w2, b2, w3, b3, w4, b4} of each network.  The SMs are
 * divided between the two weight sets in proportion to their boards and every CTA gets the same share, so a
 * batch of 5 000 runs as 104 boards on each of 146 SMs instead of three 125-CTA launches queueing for the
 * same SMs.  Results are bit-identical to the separate calls. */
int qnet_conv_forward_update_f64(const double* states, const double* next_states, const double* const* online,
                                 const double* const* target, double* q, double* patches2, double* act2,
                                 double* act3, double* q_next_online, double* q_next_target, int64_t n,
                                 void* stream);

/* ---- weight + bias gradient of a layer with a tiny weight matrix ----------------------------------- */

/* In train_step's backward (src/dqn_lib.py:159-161): dW[c][k] = sum_r g[r][c] * x[r][k] and
 * db[c] = sum_r g[r][c] for the first convolution (as a GEMM over patches, weight 64 x 4) and the
 * output layer (Linear(64,4)) of src/configs/double_dqn_conv.py:19-28 — a tall-skinny reduction over
 * 45 000 / 5 000 rows that cuBLAS + ATen run on one or two CTAs.  g [rows,C], x [rows,K] row-major
 * float64, C, K <= 64, C*K <= 1024; dw [C,K], db [C]; fixed summation order (bit-reproducible).
 * scratch: layer_wgrad_small_scratch_elems(rows, C, K) doubles (for the current, initialised device; 0 on
 * bad arguments). */
int64_t layer_wgrad_small_scratch_elems(int64_t rows, int C, int K);
int layer_wgrad_small_f64(const double* g, const double* x, double* dw, double* db, double* scratch,
                          int64_t rows, int C, int K, void* stream);

/* The same gradients for a layer with 64 outputs and K = 32, 64, ... 256 inputs (the second convolution
 * as a GEMM over patches, 20 000 rows, and Linear(256,64), 5 000 rows, of the conv Q-network) on the
 * FP64 tensor cores: the rows are split over the SMs, every CTA multiplies its range with DMMA from
 * cp.async-staged shared-memory tiles, the per-CTA results are added in a fixed order.  g [rows,64],
 * x [rows,K] row-major float64, 16-byte aligned; dw [64,K], db [64].
 * scratch: layer_wgrad64_scratch_elems(rows, K) doubles, 16-byte aligned. */
int64_t layer_wgrad64_scratch_elems(int64_t rows, int K);
int layer_wgrad64_f64(const double* g, const double* x, double* dw, double* db, double* scratch, int64_t rows,
                      int K, void* stream);

/* Backward of the conv Q-network's first convolution in one pass (train_step, src/dqn_lib.py:159-161):
 * given gpatches2 = d loss / d patches2 [4n,256] and the forward patches2 (qnet_conv_forward_train_f64),
 * col2im, the ReLU mask of conv1 and dW1 [64,1,2,2] / db1 [64] against the boards' cells (states [n,16])
 * without materialising the conv1 gradient.  Fixed summation order.
 * scratch: conv1_wgrad_fused_scratch_elems(n) doubles. */
int64_t conv1_wgrad_fused_scratch_elems(int64_t n);
int conv1_wgrad_fused_f64(const double* gpatches2, const double* patches2, const double* states, double* dw1,
                          double* db1, double* scratch, int64_t n, void* stream);

/* The same backward straight from g2 = d loss / d (conv2 pre-activation) [4n,64] (rows (board, position)): the
 * patch-matrix gradient g2 w2 (w2 = conv2.weight as [64,256]) stays in the tensor-core accumulators and is
 * consumed in place — mask, dW1, db1 — so neither it nor the conv1 gradient ever exists in memory.  One DMMA
 * kernel + the fixed-order sum of its per-CTA partial results.  g2, patches2 16-byte aligned.
 * scratch: conv2_dgrad_conv1_wgrad_scratch_elems(n) doubles. */
int64_t conv2_dgrad_conv1_wgrad_scratch_elems(int64_t n);
int conv2_dgrad_conv1_wgrad_f64(const double* g2, const double* w2, const double* patches2, const double* states,
                                double* dw1, double* db1, double* scratch, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2048_H_ */
