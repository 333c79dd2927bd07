/*
 * board_oracle.c — CPU restatement of the reference's Board2048 environment step.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
 * call it.  Nothing under reinforcement-learning-2048_b200/ imports it; the product has no CPU
 * fallback.
 *
 * Each function follows the reference (ribal-aladeeb/reinforcement-learning-2048, paths relative
 * to its root) and cites the lines it restates.  Parity is PINNED: tests/test_oracle_golden.py
 * checks this file against the .npz fixtures under tests/golden/, which oracle/gen_golden.py produced by running the
 * reference's own src/board.py (all 65536 rows, thousands of boards x 4 actions, full games) plus
 * the 18 known-answer vectors of the reference's tests/test_game_board.py.
 *
 * Boards here are int64 tile VALUES [16] row-major exactly like the reference's `state`; the
 * packed-u64 helpers at the bottom mirror include/b2048.h so that tests can feed both sides the
 * same bits.
 */
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define K 4

/* src/board.py:92-126  Board2048._apply_action_to_vector: slide one length-4 vector toward index
 * 0.  A literal restatement of the reference's while-loop (a `current` cursor and the list of
 * non-zero indices beyond it), NOT the compress-merge-compress shortcut the CUDA table builder
 * uses — so that agreement between the two is evidence, not tautology.
 * `*mergescore` accumulates like self._mergescore (src/board.py:114). */
static void ref_apply_action_to_vector(const int64_t* in, int64_t* v, int64_t* mergescore) {
  memcpy(v, in, K * sizeof(int64_t));
  int current = 0;
  while (current < K - 1) {
    /* non_zero_indices = np.where(vector != 0)[0] */
    int last_nz = -1, first_beyond = -1;
    for (int i = 0; i < K; ++i)
      if (v[i] != 0) {
        last_nz = i;
        if (i > current && first_beyond < 0) first_beyond = i;
      }
    /* if len(nz) == 0 or nz[-1] <= current: return */
    if (last_nz < 0 || last_nz <= current) return;
    /* nz = nz[current < nz]; if len(nz) == 0: return */
    if (first_beyond < 0) return;
    if (v[current] == 0) {
      v[current] += v[first_beyond];
      v[first_beyond] = 0;
    } else if (v[current] == v[first_beyond]) {
      v[current] += v[first_beyond];
      *mergescore += v[current];
      v[first_beyond] = 0;
      current += 1;
    } else if (current + 1 == first_beyond) {
      current += 1;
    } else {
      v[current + 1] = v[first_beyond];
      v[first_beyond] = 0;
      current += 1;
    }
  }
}

/* src/board.py:147-183  up/down/left/right WITHOUT the spawn: rows of state (left/right) or of
 * state.T (up/down), reversed before and after for right/down.  action: 0 up, 1 down, 2 left,
 * 3 right (src/board.py:129,191).  Returns 1 iff the board changed (src/board.py:151). */
int oracle_slide(const int64_t* in, int action, int64_t* out, int64_t* reward) {
  int64_t score = 0;
  for (int line = 0; line < K; ++line) {
    int64_t vec[K], res[K];
    for (int i = 0; i < K; ++i) {
      const int j = (action == 1 || action == 3) ? (K - 1 - i) : i; /* _reverse_vector */
      vec[i] = (action < 2) ? in[j * K + line]   /* column `line` = row of state.T */
                            : in[line * K + j];  /* row `line` */
    }
    ref_apply_action_to_vector(vec, res, &score);
    for (int i = 0; i < K; ++i) {
      const int j = (action == 1 || action == 3) ? (K - 1 - i) : i;
      if (action < 2) out[j * K + line] = res[i];
      else out[line * K + j] = res[i];
    }
  }
  *reward = score;
  return memcmp(in, out, K * K * sizeof(int64_t)) != 0;
}

/* src/board.py:128-135  available_moves_as_torch_unit_vector: move legal iff it changes the board.
 * Bits: 1 up, 2 down, 4 left, 8 right. */
int oracle_legal_mask(const int64_t* in) {
  int m = 0;
  for (int a = 0; a < 4; ++a) {
    int64_t out[K * K], r;
    if (oracle_slide(in, a, out, &r)) m |= 1 << a;
  }
  return m;
}

/* src/board.py:41-51  _populate_empty_cell with the random draws supplied by the caller:
 * `rank` = random.randint(0, n_empty-1) into the row-major list of empty cells (np.where order),
 * `value` = the drawn 2 or 4.  Returns the cell index or -1. */
int oracle_populate(int64_t* state, int rank, int64_t value) {
  int seen = 0;
  for (int i = 0; i < K * K; ++i)
    if (state[i] == 0) {
      if (seen == rank) {
        state[i] = value;
        return i;
      }
      ++seen;
    }
  return -1;
}

static int count_empty(const int64_t* s) {
  int n = 0;
  for (int i = 0; i < K * K; ++i) n += (s[i] == 0);
  return n;
}

/* flags byte layout of include/b2048.h */
enum { F_DONE = 0x10, F_CHANGED = 0x20, F_OVERFLOW = 0x40, F_BADSPAWN = 0x80 };

/* One environment step on tile values: src/board.py:185-202 peek_action + src/dqn_lib.py:17-18
 * (done = no legal move on the INPUT board) + src/dqn_lib.py:87-88 (reward = merge-score delta).
 * spawn_cell < 0 means "no spawn information": the slid board is returned without a new tile. */
int oracle_step_tiles(const int64_t* in, int action, int spawn_cell, int64_t spawn_value,
                      int64_t* out, int64_t* reward) {
  const int legal = oracle_legal_mask(in);
  int flags = legal | (legal ? 0 : F_DONE);
  const int changed = oracle_slide(in, action & 3, out, reward);
  if (changed) {
    flags |= F_CHANGED;
    if (spawn_cell >= 0) {
      if (out[spawn_cell] != 0) flags |= F_BADSPAWN;
      else out[spawn_cell] = spawn_value;
    }
  }
  return flags;
}

/* ---- packed-u64 mirror of include/b2048.h ----------------------------------------------------- */

static int exp_of(int64_t t) {
  int e = 0;
  while (t > 1) {
    t >>= 1;
    ++e;
  }
  return e;
}

uint64_t oracle_pack(const int64_t* tiles) {
  uint64_t b = 0;
  for (int i = 0; i < 16; ++i) b |= (uint64_t)(tiles[i] ? exp_of(tiles[i]) : 0) << (4 * i);
  return b;
}

void oracle_unpack(uint64_t b, int64_t* tiles) {
  for (int i = 0; i < 16; ++i) {
    const int e = (int)((b >> (4 * i)) & 0xF);
    tiles[i] = e ? ((int64_t)1 << e) : 0;
  }
}

/* Philox4x32-R, restated from Salmon et al. SC'11 (Random123 philox.h constants), to let the
 * oracle predict the library's streams.  R = 10 is checked against the Random123 known-answer vectors
 * in tests/test_oracle_golden.py; the spawn stream uses R = 7 (Random123's philox4x32_7), all other
 * streams R = 10. */
void oracle_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4], int rounds) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int i = 0; i < rounds; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  oracle_philox4x32_r(ctr, key, out, 10);
}

#define SPAWN_PHILOX_ROUNDS 7
#define DOM_SPAWN 0x00000000u
#define DOM_RESET 0x5BD1E995u

/* The library's spawn stream, ABI version 2 (include/b2048.h): board g owns the 16-bit lane (g & 7) of
 * the Philox4x32-7 call with counter (g >> 3, step) and key (seed ^ domain); lane j is half (j & 1)
 * of output word (j >> 1), low half first.  Returns the lane value d, 0..65535. */
uint32_t oracle_spawn_draw(uint64_t seed, uint64_t step, uint64_t g) {
  const uint64_t pidx = g >> 3;
  const uint32_t ctr[4] = {(uint32_t)pidx, (uint32_t)(pidx >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ DOM_SPAWN};
  uint32_t o[4];
  oracle_philox4x32_r(ctr, key, o, SPAWN_PHILOX_ROUNDS);
  const uint32_t lane = (uint32_t)(g & 7u), w = o[lane >> 1];
  return (lane & 1u) ? (w >> 16) : (w & 0xFFFFu);
}

/* Batched step on packed boards with the library's documented spawn rule (include/b2048.h):
 * with d = the board's 16-bit draw and n = number of empty cells of the slid board,
 * cell = k-th empty (row-major), k = floor(d * n / 65536); "4" iff (((d * n) mod 65536) << 16) <
 * p4_threshold; spawn_override[i] != 0xFF replays (cell | exp << 4). */
static void step_packed_range(const uint64_t* boards, const uint8_t* actions, uint64_t* next,
                              int32_t* reward, uint8_t* flags, int64_t begin, int64_t end,
                              uint64_t seed, uint64_t step, uint64_t index_base,
                              uint32_t p4_threshold, const uint8_t* spawn_override) {
  for (int64_t i = begin; i < end; ++i) {
    int64_t in[16], out[16], r = 0;
    oracle_unpack(boards[i], in);
    int cell = -1;
    int64_t val = 0;
    int f;
    if (spawn_override && spawn_override[i] == 0xFE) {          /* B2048_SPAWN_SKIP: slide-only result */
      f = oracle_step_tiles(in, actions[i] & 3, -1, 0, out, &r);
    } else if (spawn_override && spawn_override[i] != 0xFF) {
      cell = spawn_override[i] & 0xF;
      val = (int64_t)1 << (spawn_override[i] >> 4);
      f = oracle_step_tiles(in, actions[i] & 3, cell, val, out, &r);
    } else {
      f = oracle_step_tiles(in, actions[i] & 3, -1, 0, out, &r);
      if (f & F_CHANGED) {
        const uint32_t d = oracle_spawn_draw(seed, step, index_base + (uint64_t)i);
        const uint32_t ne = (uint32_t)count_empty(out);
        const uint32_t prod = d * ne;                       /* < 2^20 */
        const int rank = (int)(prod >> 16);
        const uint32_t frac = (prod & 0xFFFFu) << 16;
        oracle_populate(out, rank, (frac < p4_threshold) ? 4 : 2);
      }
    }
    /* 32768 + 32768 does not fit a nibble: the library flags it and leaves next/reward unspecified */
    for (int c = 0; c < 16; ++c)
      if (out[c] > 32768) f |= F_OVERFLOW;
    next[i] = oracle_pack(out);
    reward[i] = (int32_t)r;
    flags[i] = (uint8_t)f;
  }
}

/* The library's synthetic-input generators (b2048_random_boards / b2048_random_actions, include/b2048.h),
 * restated so that the bench stream exists on a CPU-only box too: oracle/gen_golden.py feeds its first
 * 65536 boards to the reference, and __graft_entry__.smoke() checks the kernel against those goldens.
 * Board g: 8 Philox calls (counter (g, q), q = 0..7, domain BOARDS), words (x,y) and (z,w) give cells
 * 2q and 2q+1: empty iff the first word < p_empty_threshold, else exponent 1 + floor(second * max_exp / 2^32).
 * Action g: word (g & 3) of the call with counter (g >> 2, step), domain ACTIONS; action = word >> 30. */
#define DOM_BOARDS 0x1B873593u
#define DOM_ACTIONS 0xCC9E2D51u
void oracle_stream_boards(uint64_t* boards, int64_t n, uint64_t seed, uint64_t index_base,
                          uint32_t p_empty_threshold, uint32_t max_exp) {
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ DOM_BOARDS};
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t g = index_base + (uint64_t)i;
    uint64_t b = 0;
    for (uint32_t q = 0; q < 8; ++q) {
      const uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), q, 0u};
      uint32_t o[4];
      oracle_philox4x32_10(ctr, key, o);
      const uint64_t e0 = (o[0] < p_empty_threshold) ? 0u : 1u + (uint32_t)(((uint64_t)o[1] * max_exp) >> 32);
      const uint64_t e1 = (o[2] < p_empty_threshold) ? 0u : 1u + (uint32_t)(((uint64_t)o[3] * max_exp) >> 32);
      b |= e0 << (8 * q);
      b |= e1 << (8 * q + 4);
    }
    boards[i] = b;
  }
}

void oracle_stream_actions(uint8_t* actions, int64_t n, uint64_t seed, uint64_t step, uint64_t index_base) {
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ DOM_ACTIONS};
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t g = index_base + (uint64_t)i, pidx = g >> 2;
    const uint32_t ctr[4] = {(uint32_t)pidx, (uint32_t)(pidx >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    actions[i] = (uint8_t)(o[g & 3] >> 30);
  }
}

/* Fresh boards with the library's reset stream: src/board.py:10-20 (zeros + two spawns). */
void oracle_reset_packed(uint64_t* boards, int64_t n, uint64_t seed, uint64_t step,
                         uint64_t index_base, uint32_t p4_threshold) {
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t g = index_base + (uint64_t)i;
    const uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ DOM_RESET};
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    int64_t t[16];
    memset(t, 0, sizeof t);
    t[o[0] >> 28] = (o[1] < p4_threshold) ? 4 : 2;
    const int rank = (int)(((uint64_t)o[2] * 15u) >> 32);
    oracle_populate(t, rank, (o[3] < p4_threshold) ? 4 : 2);
    boards[i] = oracle_pack(t);
  }
}

void oracle_legal_mask_packed(const uint64_t* boards, uint8_t* flags, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    int64_t in[16];
    oracle_unpack(boards[i], in);
    const int m = oracle_legal_mask(in);
    flags[i] = (uint8_t)(m | (m ? 0 : F_DONE));
  }
}

/* One row through the reference's vector routine (for the 65536-row golden check). */
void oracle_row_left(const int64_t* in4, int64_t* out4, int64_t* reward) {
  int64_t s = 0;
  ref_apply_action_to_vector(in4, out4, &s);
  *reward = s;
}

int oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

typedef struct {
  const uint64_t* boards; const uint8_t* actions; uint64_t* next; int32_t* reward; uint8_t* flags;
  int64_t begin, end; uint64_t seed, step, index_base; uint32_t p4; const uint8_t* ovr;
} step_job;

static void* step_worker(void* p) {
  step_job* j = (step_job*)p;
  step_packed_range(j->boards, j->actions, j->next, j->reward, j->flags, j->begin, j->end, j->seed,
                    j->step, j->index_base, j->p4, j->ovr);
  return 0;
}

/* `threads` <= 1 runs inline; otherwise the batch is split into contiguous ranges, one pthread
 * each (this is the "port" CPU baseline of bench.py: all host cores). */
void oracle_step_packed_mt(const uint64_t* boards, const uint8_t* actions, uint64_t* next,
                           int32_t* reward, uint8_t* flags, int64_t n, uint64_t seed, uint64_t step,
                           uint64_t index_base, uint32_t p4_threshold, const uint8_t* spawn_override,
                           int threads) {
  if (threads <= 1 || n < 1024) {
    step_packed_range(boards, actions, next, reward, flags, 0, n, seed, step, index_base,
                      p4_threshold, spawn_override);
    return;
  }
  if (threads > 256) threads = 256;
  pthread_t tid[256];
  step_job jobs[256];
  const int64_t per = (n + threads - 1) / threads;
  int started = 0;
  for (int t = 0; t < threads; ++t) {
    const int64_t b = per * t, e = (b + per < n) ? b + per : n;
    if (b >= e) break;
    jobs[t] = (step_job){boards, actions, next, reward, flags, b, e, seed, step, index_base,
                         p4_threshold, spawn_override};
    pthread_create(&tid[t], 0, step_worker, &jobs[t]);
    ++started;
  }
  for (int t = 0; t < started; ++t) pthread_join(tid[t], 0);
}

void oracle_step_packed(const uint64_t* boards, const uint8_t* actions, uint64_t* next,
                        int32_t* reward, uint8_t* flags, int64_t n, uint64_t seed, uint64_t step,
                        uint64_t index_base, uint32_t p4_threshold, const uint8_t* spawn_override) {
  step_packed_range(boards, actions, next, reward, flags, 0, n, seed, step, index_base,
                    p4_threshold, spawn_override);
}
