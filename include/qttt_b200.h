/* qttt_b200.h -- C ABI of libqttt_b200.so: batched quantum tic-tac-toe transitions on B200.
 *
 * Every entry point replaces a piece of the reference's (Oxel40/qtttgym) game-transition
 * path; the citation after each one names the reference interface it stands in for (paths
 * relative to the reference checkout).  The reference has no FFI of its own -- it is pure
 * Python -- so "what its FFI would bind" is: the gym-like Env API (qtttgym/env.py:15-66),
 * the duck-typed evaluator seam Board(qevaluator).eval (qtttgym/board.py:2,7,51), and the
 * GameState / rollout helpers MCTS relies on (mcts.py:19-27,52-65,87-91,185-208,233-267).
 * INTEGRATION.md shows the ctypes stub a qtttgym maintainer would add.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.  The caller owns every
 *     buffer; the library allocates nothing and keeps no mutable global state.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous and stream-ordered; they are re-entrant across streams and devices.
 *   - Return value: QTTT_OK, a negative QTTT_ERR_* code, or -(1000 + cudaError_t).
 *     Nothing throws.  qttt_strerror() never returns NULL.
 *   - Optional outputs may be NULL.
 *
 * Packed game state (qttt_state, 16 bytes, opaque to callers; see csrc/qttt_core.cuh):
 * holds what the reference keeps in Board.moves / Board.board (qtttgym/board.py:4-5);
 * Board.qstructs (board.py:6) is derived on the fly.  Convert with qttt_pack/qttt_observe.
 *
 * Action encodings
 *   QTTT_ACT_INDEX  uint8[n]     0..35, pair (i<j) in lexicographic order (mcts.py:339-349);
 *                                any other value is an illegal action (no-op).
 *   QTTT_ACT_PAIR   int8[n][2]   (a, b) exactly as passed to Env.step (env.py:34-40), any
 *                                order; a == b, a classical square, or a value outside 0..8
 *                                is an illegal action: state unchanged, turn not advanced,
 *                                outputs recomputed (env.py:36-43).
 */
#ifndef QTTT_B200_H
#define QTTT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTTT_ABI_VERSION 2

#if defined(__GNUC__)
#define QTTT_API __attribute__((visibility("default")))
#else
#define QTTT_API
#endif

#define QTTT_OK 0
#define QTTT_ERR_ARG (-1)        /* NULL / negative size / bad enum */
#define QTTT_ERR_ALIGN (-2)      /* a buffer is not aligned for its element type */
#define QTTT_ERR_NO_DEVICE (-3)  /* no CUDA device or not an sm_100 device */

#define QTTT_ACT_INDEX 0
#define QTTT_ACT_PAIR 1

/* per-game status byte written by qttt_step* */
#define QTTT_ST_OK 0        /* move accepted */
#define QTTT_ST_ILLEGAL 1   /* the reference would have raised; swallowed no-op (env.py:41-43) */
#define QTTT_ST_FINISHED 2  /* random-policy step on a terminated game: nothing to do */
#define QTTT_ST_RESET 4     /* OR-ed in: the game was over on entry and was reset first (autoreset) */

/* flags of qttt_step_ex / qttt_step_random_ex (at most one) */
#define QTTT_STEP_FRESH 1u           /* Env.reset then Env.step: the incoming state is ignored */
#define QTTT_STEP_AUTORESET 2u       /* a game that is over on entry is reset, then the action is applied */
#define QTTT_STEP_AUTORESET_NEXT 4u  /* ... is reset and its action ignored (vector-env "next step" autoreset) */

typedef struct { uint32_t w[4]; } qttt_state;

QTTT_API int qttt_abi_version(void);
QTTT_API const char* qttt_strerror(int rc);

/* Board.__init__ / Env.reset (qtttgym/board.py:2-7, qtttgym/env.py:55-57; the seed is
 * ignored there, Q4).  Writes n empty games; mask (optional) gets the 36-bit legal mask of
 * the empty board (all 36 actions). */
QTTT_API int qttt_reset(qttt_state* state, uint64_t* mask, int64_t n, void* stream);

/* qttt_reset that also writes what Env.step would report for a fresh game (optional outputs):
 * reward -0.0f (env.py:49 with no line), done 0, status QTTT_ST_OK. */
QTTT_API int qttt_reset_all(qttt_state* state, uint64_t* mask, float* reward, uint8_t* done,
                            uint8_t* status, int64_t n, void* stream);

/* Env.step (qtttgym/env.py:34-53) for n independent games, including Board.make_move /
 * update_qstructs (board.py:9-69), QEvalClassic.eval (qeval.py:5-51) and check_win
 * (board.py:71-115).
 *   action, action_format : see above
 *   coin   uint8[n]       : forced measurement outcome, consumed only by games whose move
 *                           closes a cycle: 0 -> the closing move falls into its smaller
 *                           square (index random.choice picks at qeval.py:35).  NULL = draw it
 *                           from Philox4x32-10 keyed (seed, game_base + i, len(moves)).
 *   reward float[n]       : env.py:49 bit-exact: -1.0f if any line exists else -0.0f
 *   done   uint8[n]       : env.py:51 terminated
 *   mask   uint64[n]      : 36-bit legal-action mask of the new state (mcts.py:87-91)
 *   status uint8[n]       : QTTT_ST_* */
QTTT_API int qttt_step(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
              uint64_t seed, uint64_t game_base, float* reward, uint8_t* done, uint64_t* mask,
              uint8_t* status, int64_t n, void* stream);

/* Env.reset followed by the first Env.step (env.py:55-57 then 34-53) in one launch: the games
 * start from the empty board, so the state is written but not read.  Same arguments and
 * outputs as qttt_step. */
QTTT_API int qttt_reset_step(qttt_state* state, const void* action, int action_format,
                             const uint8_t* coin, uint64_t seed, uint64_t game_base, float* reward,
                             uint8_t* done, uint64_t* mask, uint8_t* status, int64_t n, void* stream);

/* qttt_step with an episode counter and a mode.
 *   epoch : folded into the Philox counter of the collapse coin (coin == NULL): the coin of game
 *           g at ply p in epoch e is a function of (seed, g, p, e).  The reference draws a fresh
 *           random.choice per collapse (qeval.py:35); a caller that plays episode after episode
 *           in the same env slots passes a different epoch per episode (qtttgym_b200.BatchedEnv
 *           bumps it in reset()) or the episodes would all see the same coins.  epoch 0 is the
 *           stream of qttt_step.  (24 bits are used.)
 *   flags : 0, or one of QTTT_STEP_FRESH (== qttt_reset_step), QTTT_STEP_AUTORESET,
 *           QTTT_STEP_AUTORESET_NEXT.  With an autoreset flag a game that is over on entry
 *           (a line exists or len(moves) == 9: what Env.step reports as terminated, env.py:51)
 *           restarts from the empty board inside the same launch (status gets QTTT_ST_RESET),
 *           so a batch driven by a policy never idles and never needs a host round trip. */
QTTT_API int qttt_step_ex(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
                          uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags,
                          float* reward, uint8_t* done, uint64_t* mask, uint8_t* status, int64_t n,
                          void* stream);

/* qttt_step with compact I/O, for callers whose buffers live in HOST memory (there the PCIe
 * link, not HBM, is the bound: 3 bytes per game cross it instead of 15).
 *   action_coin uint8[n]  : bits 0..5 action index (QTTT_ACT_INDEX; 36..63 = illegal),
 *                           bit 7 forced coin
 *   result      uint16[n] : bits 0..8 the free (non-classical) squares of the new state -- the
 *                           36-bit legal mask is a function of it (legal[k] = both squares of
 *                           pair k free, mcts.py:19-27; qtttgym_b200.unpack_result expands it
 *                           with a 512-entry table) --, bit 9 terminated, bit 10 "a line
 *                           exists" (reward = bit ? -1.0f : -0.0f, env.py:49), bits 11..12
 *                           status (QTTT_ST_*)
 * Same transition, same information as qttt_step, bit for bit. */
QTTT_API int qttt_step_packed(qttt_state* state, const uint8_t* action_coin, uint16_t* result,
                              int64_t n, void* stream);

/* qttt_step_packed that also writes the observation: obs qttt_state[n] (optional) receives the
 * post-step packed state of every game (for an illegal no-op: the unchanged state), i.e. what
 * Env.step returns as obs (env.py:46,68-85) in packed form. */
QTTT_API int qttt_step_packed_obs(qttt_state* state, const uint8_t* action_coin, uint16_t* result,
                                  qttt_state* obs, int64_t n, void* stream);

/* qttt_step_packed_obs for MAPPED PINNED HOST buffers (cudaHostAlloc / cudaHostRegister memory,
 * device-accessible under unified addressing): the kernel itself reads action_coin_host and
 * writes result_host / obs_host (optional) across PCIe in 512-byte bursts -- no copy engine, no
 * device staging buffers, one launch.  The caller synchronises the stream before reading. */
QTTT_API int qttt_step_packed_mapped(qttt_state* state, const uint8_t* action_coin_host,
                                     uint16_t* result_host, qttt_state* obs_host, int64_t n, void* stream);

/* qttt_step_packed_mapped with the results bit-packed: a result word carries 12 bits (free-square
 * set 0-8, terminated 9, line 10, illegal 11), so four consecutive games' results are written as
 * THREE 16-bit words -- word k (k = 0..2) of group g = games 4g..4g+3 holds game 4g+k's result in
 * bits 0-11 and bits 4k..4k+3 of game 4g+3's result in bits 12-15.  result12_host:
 * uint16[3 * ceil(n / 4)] (games past n read as 0).  1.5 instead of 2 bytes per game cross the
 * link, which is what bounds the host-resident caller. */
QTTT_API int qttt_step_packed12_mapped(qttt_state* state, const uint8_t* action_coin_host,
                                       uint16_t* result12_host, int64_t n, void* stream);

/* The bit-packed results through explicit copies: qttt_step_packed_host with 12-bit results.  Slices
 * of `slice` games (a multiple of 4) are pipelined over `streams`: cudaMemcpyAsync of the slice's
 * action bytes into in_dev[n], the step kernel writing packed words into out12_dev
 * (uint16[3 * ceil(n / 4)]), cudaMemcpyAsync of those words into result12_host. */
QTTT_API int qttt_step_packed12_host(qttt_state* state, const uint8_t* action_coin_host,
                                     uint16_t* result12_host, uint8_t* in_dev, uint16_t* out12_dev,
                                     int64_t n, int64_t slice, void* const* streams, int n_streams);

/* Env.step for a host-resident caller that wants the OBSERVATION back (env.py:34-53,68-85) at 12
 * bytes per game instead of 18 (16-byte packed state + result word): after the step the kernel
 * writes one record of three 32-bit words per game into obs12_dev (uint32[n][3], device staging),
 * and each slice is copied to obs12_host (pinned) -- pipelined over `streams` like
 * qttt_step_packed_host.
 *   word 0  classical squares 0..7, 4 bits each: owning move index + 1, 0 = not classical
 *   word 1  square 8 (bits 0-3) | move slots 0..3, 6 bits each from bit 4 |
 *           terminated (28) | line, i.e. reward -1.0 (29) | illegal no-op (30)
 *   word 2  move slots 4..8, 6 bits each
 * A slot holds the action index (0..35, mcts.py:339-349) of an UNCOLLAPSED move and 63 otherwise:
 * the even slots in order are obs["q_states_p1"], the odd ones obs["q_states_p2"];
 * obs["turn"] = (classical squares + uncollapsed moves) % 2; the legal mask follows from the free
 * squares. */
QTTT_API int qttt_step_packed_host_obs12(qttt_state* state, const uint8_t* action_coin_host,
                                         uint32_t* obs12_host, uint8_t* in_dev, uint32_t* obs12_dev,
                                         int64_t n, int64_t slice, void* const* streams, int n_streams);

/* The host-buffer form of qttt_step_packed: action_coin_host / result_host are PINNED HOST
 * arrays; in_dev (uint8[n]) / out_dev (uint16[n]) are caller-provided device staging buffers.
 * The batch is cut into slices of `slice` games; slice k is copied in, stepped and copied out
 * on streams[k % n_streams], so that host->device copies, kernels and device->host copies of
 * different slices overlap.  The caller orders the streams against its own work. */
QTTT_API int qttt_step_packed_host(qttt_state* state, const uint8_t* action_coin_host,
                                   uint16_t* result_host, uint8_t* in_dev, uint16_t* out_dev,
                                   int64_t n, int64_t slice, void* const* streams, int n_streams);

/* qttt_step_packed_host that also brings the observation back: obs_host qttt_state[n] (pinned,
 * optional) receives each slice of the state array after its kernel (16 more bytes per game
 * over PCIe). */
QTTT_API int qttt_step_packed_host_obs(qttt_state* state, const uint8_t* action_coin_host,
                                       uint16_t* result_host, qttt_state* obs_host, uint8_t* in_dev,
                                       uint16_t* out_dev, int64_t n, int64_t slice, void* const* streams,
                                       int n_streams);

/* Env.step driven by the uniform-random policy of MCTS._simulate (mcts.py:185-198,
 * 287-292): action ~ U(legal actions), coin ~ U{0,1}, both from Philox4x32-10 with counter
 * (game_lo, game_hi, len(moves), 0) and key seed.  Terminated games (mcts.py:52-65) are left
 * untouched with status QTTT_ST_FINISHED.  action_out / coin_out (optional, uint8[n]) record
 * the trace (255 / 0 for untouched games). */
QTTT_API int qttt_step_random(qttt_state* state, uint64_t seed, uint64_t game_base, uint8_t* action_out,
                     uint8_t* coin_out, float* reward, uint8_t* done, uint64_t* mask,
                     uint8_t* status, int64_t n, void* stream);

/* qttt_step_random with an epoch and a mode (see qttt_step_ex).  With QTTT_STEP_AUTORESET every
 * call plays one ply in every game, finished games restarting from the empty board: continuous
 * random self-play through the step API. */
QTTT_API int qttt_step_random_ex(qttt_state* state, uint64_t seed, uint64_t game_base, uint64_t epoch,
                                 uint32_t flags, uint8_t* action_out, uint8_t* coin_out, float* reward,
                                 uint8_t* done, uint64_t* mask, uint8_t* status, int64_t n, void* stream);

/* Env._observation / Env.turn / Env._reward / check_win / update_winner / action_mask in
 * tensor form (qtttgym/env.py:62-112, board.py:71-115, mcts.py:52-65,87-91).
 *   classical int8[n][9]    Board.board (-1 or owning move index)
 *   moves     int8[n][9][2] Board.moves (a, b), rows >= n_moves are (-1,-1); idx = row
 *   n_moves   uint8[n]      len(Board.moves)  (Env.turn)
 *   q_p1      int8[n][5][2] obs["q_states_p1"], padded with (-1,-1)
 *   q_p2      int8[n][4][2] obs["q_states_p2"]
 *   turn      uint8[n]      obs["turn"] = len(moves) % 2
 *   rounds    int8[n][2]    check_win() -> (p1_round, p2_round)
 *   reward_p1 float[n]      Env._reward()
 *   winner    uint8[n]      0 none/draw, 1 X, 2 O (mcts.py:52-65)
 *   mask_bool uint8[n][36]  GameState.action_mask()
 * Alignment: q_p2 8 bytes, reward_p1 4, rounds 2 (QTTT_ERR_ALIGN otherwise); the other arrays any,
 * 16 bytes for the fast path.  Three specialisations are dispatched on which outputs are non-NULL:
 * exactly {classical, q_p1, q_p2, turn} (the env.py observation), every output, anything else. */
QTTT_API int qttt_observe(const qttt_state* state, int8_t* classical, int8_t* moves, uint8_t* n_moves,
                 int8_t* q_p1, int8_t* q_p2, uint8_t* turn, int8_t* rounds, float* reward_p1,
                 uint8_t* winner, uint8_t* mask_bool, int64_t n, void* stream);

/* GameState.to_vector (mcts.py:67-85): the (18, 10) float feature matrix the reference feeds
 * its policy/value net, for n games -> features float[n][18][10] (16-byte aligned).
 * Rows 0..8: one-hot of board[square] (column 9 = not classical); rows 9..17: 1/sqrt(9) at
 * [square, t] for every move t on that square, 1.0 in column 9 for squares in no entangled
 * component.  (nn.Model.get_mask, nn.py:44-61, is the complement of the legal mask.) */
QTTT_API int qttt_features(const qttt_state* state, float* features, int64_t n, void* stream);

/* qtttgym.Env for ONE game at minimum latency (the single-env adapter qtttgym_b200.Env): one
 * launch applies `op` to state[0] and writes a 128-byte record of everything Env.step / observ /
 * turn / _reward / check_win report into record_host, which must be MAPPED PINNED HOST memory
 * (16-byte aligned); the 32-bit word at byte 124 is set to `seq` last, after a system-scope
 * fence, so the host can spin on it instead of synchronising a stream.
 *   op 0 : Env.step((a, b)) (env.py:34-53); coin < 0 -> Philox(seed, game 0, len(moves), epoch)
 *   op 1 : Env.reset (env.py:55-57)         op 2 : re-emit the record of the current state
 * Record layout (bytes): state 0..15 | legal mask u64 16 | reward f32 24 | terminated 28 |
 * status 29 | turn 30 | len(moves) 31 | classical i8[9] 32 | q_states_p1 i8[5][2] 48 |
 * q_states_p2 i8[4][2] 58 | check_win rounds i8[2] 66 | winner 68 | Env._reward f32 72 |
 * moves i8[9][2] 80 | seq u32 124. */
QTTT_API int qttt_env1(qttt_state* state, int op, int a, int b, int coin, uint64_t seed, uint64_t epoch,
                       void* record_host, uint32_t seq, void* stream);

/* QEvalClassic.eval (qeval.py:5-51) for ONE measurement at minimum latency -- the evaluator
 * plugin Board(qevaluator) calls at board.py:51.  state_host is read on the HOST (the packed
 * position travels as kernel arguments); record_host is 32 bytes of MAPPED PINNED HOST memory
 * (16-byte aligned): sq0 int8[9] | sq1 int8[9] (the square each move index collapses into for
 * coin 0 / 1, -1 = not measured) | closes at byte 18 | seq u32 at byte 28, written last after a
 * system-scope fence (spin on it). */
QTTT_API int qttt_qeval1(const qttt_state* state_host, int action, void* record_host, uint32_t seq,
                         void* stream);

/* nn.Model.get_mask (nn.py:44-61) for packed states: illegal_mask uint8[n][36] (bool bytes,
 * 4-byte aligned), entry [g][a] = 1 when action a touches a classical square of game g
 * (occupied[i] or occupied[j]) -- the logits the reference's policy head sets to -inf.  It is
 * the complement of GameState.action_mask() (mcts.py:87-91). */
QTTT_API int qttt_get_mask(const qttt_state* state, uint8_t* illegal_mask, int64_t n, void* stream);

/* qttt_step_ex fused with the net-input encoding of the NEW state: features float[n][18][10]
 * (GameState.to_vector, mcts.py:67-85; required, 16-byte aligned) and, optionally,
 * illegal_mask uint8[n][36] (nn.Model.get_mask, nn.py:44-61) are written by the same launch that
 * steps the games, so feeding a policy/value net costs no second pass over the state array
 * (what AlphaZero-style callers do after every move: alphazero.py:143-144,294-303).  All other
 * arguments and outputs as in qttt_step_ex; reward / done / mask / status may be NULL. */
QTTT_API int qttt_step_features(qttt_state* state, const void* action, int action_format,
                                const uint8_t* coin, uint64_t seed, uint64_t game_base, uint64_t epoch,
                                uint32_t flags, float* reward, uint8_t* done, uint64_t* mask,
                                uint8_t* status, float* features, uint8_t* illegal_mask, int64_t n,
                                void* stream);

/* qttt_step_ex fused with the env.py observation of the NEW state (what Env.step returns as obs,
 * qtttgym/env.py:34-53 with env.py:68-85): classical int8[n][9], q_p1 int8[n][5][2], q_p2
 * int8[n][4][2] (8-byte aligned), turn uint8[n] -- all four required, laid out as in qttt_observe
 * -- are written by the launch that steps the games, from the state it still holds in registers.
 * All other arguments and outputs as in qttt_step_ex; reward / done / mask / status may be NULL. */
QTTT_API int qttt_step_obs(qttt_state* state, const void* action, int action_format, const uint8_t* coin,
                           uint64_t seed, uint64_t game_base, uint64_t epoch, uint32_t flags, float* reward,
                           uint8_t* done, uint64_t* mask, uint8_t* status, int8_t* classical, int8_t* q_p1,
                           int8_t* q_p2, uint8_t* turn, int64_t n, void* stream);

/* Inverse of qttt_observe for (classical, moves, n_moves): builds packed states from
 * reference-shaped positions (what MCTS.reset does with game.board / game.moves,
 * mcts.py:139-164 -- but the entanglement is re-derived, so mid-game roots are handled
 * correctly, unlike the reference).  Positions must be reachable ones. */
QTTT_API int qttt_pack(qttt_state* state, const int8_t* classical, const int8_t* moves,
              const uint8_t* n_moves, int64_t n, void* stream);

/* The measurement seam: Board.update_qstructs -> QEvalClassic.eval (board.py:42-56,
 * qeval.py:5-51) and MCTS._step's enumeration of BOTH collapse outcomes (mcts.py:233-267).
 * For each game and action (QTTT_ACT_INDEX) computes the two successor positions.
 *   next0/next1 qttt_state[n] : successor for coin 0 / 1 (identical when no cycle closes)
 *   board0/board1 uint64[n]   : successor boards, 4 bits per square, value board[s]+1
 *   sq0/sq1 int8[n][9]        : square each move index collapses into in this measurement
 *                               (-1 = not part of it): eval()'s return value by move index
 *   closes uint8[n]           : 1 if the action closes a cycle (two distinct outcomes)
 *   result_prob float[n][3]   : P(X has the earlier line), P(O ...), P(neither) over the two
 *                               equiprobable outcomes (values in {0, .5, 1}) */
QTTT_API int qttt_qeval_both(const qttt_state* state, const uint8_t* action, qttt_state* next0,
                    qttt_state* next1, uint64_t* board0, uint64_t* board1, int8_t* sq0,
                    int8_t* sq1, uint8_t* closes, float* result_prob, int64_t n, void* stream);

/* MCTS leaf evaluation (mcts.py:166-173 _rollout averaging, 185-208 _simulate/_reward):
 * n_rollouts uniform-random playouts from every root; playout j of root r uses Philox
 * game id r * n_rollouts + j, domain 1.
 *   tallies int32[n_roots][3] : X wins, O wins, draws
 *   value   float[n_roots]    : mean reward from the root's side to move (mcts.py:171,173)
 *   steps_total int64[1]      : env-steps executed inside the playouts (ADDED to) */
QTTT_API int qttt_rollout(const qttt_state* roots, int64_t n_roots, int32_t n_rollouts, uint64_t seed,
                 int32_t* tallies, float* value, int64_t* steps_total, void* stream);

/* ---- MCTS search around the leaf evaluator (mcts.py:132-337), one tree per root ------------
 * The caller provides the node pool: n_roots * capacity nodes of qttt_mcts_node_bytes() bytes
 * (16-byte aligned) and meta int32[n_roots][8].  A search of R rollouts needs at most
 * 1 + 2 * R nodes (+ 2 per sync); if the pool runs out meta[r][3] gets bit 0 set and the tree
 * stops growing.  Random choices follow the keyed Philox stream of oracle/mcts_oracle.py
 * (pinned to the unmodified reference MCTS), so the statistics are bit-identical to the
 * reference's for the same stream.  root_base + n_roots <= 2^20, num_simulations <= 4096. */
QTTT_API int qttt_mcts_node_bytes(void);
/* MCTS.reset (mcts.py:139-164) for every root. */
QTTT_API int qttt_mcts_init(void* pool, int64_t capacity, int32_t* meta, const qttt_state* roots,
                            int64_t n_roots, void* stream);
/* n_rollouts x MCTS._rollout (mcts.py:166-173): PUCT select (269-285), expansion with both
 * collapse outcomes (210-267), num_simulations random playouts of the leaf (185-208),
 * backpropagation (175-183).  c_puct = 1.0 in the reference (mcts.py:134). */
QTTT_API int qttt_mcts_run(void* pool, int64_t capacity, int32_t* meta, int32_t n_rollouts,
                           int32_t num_simulations, double c_puct, uint64_t seed, uint64_t root_base,
                           int64_t n_roots, void* stream);
/* root.N / root.Q / root.Ntot and MCTS.choose (mcts.py:308-315): n_visits int32[n][36],
 * q_values double[n][36], n_total int32[n], choose uint8[n]; any may be NULL. */
QTTT_API int qttt_mcts_stats(const void* pool, int64_t capacity, const int32_t* meta,
                             int32_t* n_visits, double* q_values, int32_t* n_total, uint8_t* choose,
                             int64_t n_roots, void* stream);
/* MCTS.sync (mcts.py:317-337): the root moves to the child of action[r] whose position is
 * now[r] (expanding it if needed); meta[r][3] bit 1 is set when there is no such child.
 * action[r] >= 36 means "no move was played in this game" (a finished game in a batch): the
 * root stays where it is and no error is flagged. */
QTTT_API int qttt_mcts_sync(void* pool, int64_t capacity, int32_t* meta, const uint8_t* action,
                            const qttt_state* now, int64_t n_roots, void* stream);

/* Random self-play sweep (strat_eval.py:66-94 tally convention): plays games with global
 * ids [game_lo, game_hi) from the empty board to termination with the random policy
 * (domain 0), entirely in registers.  stats int64[16] is ADDED to:
 *   [0] X wins [1] O wins [2] draws [3] env-steps [4] collapses [5] games
 *   [6..15] games by number of env-steps (0..9)
 * Sharding [game_lo, game_hi) over GPUs and summing stats gives the single-GPU result. */
QTTT_API int qttt_sweep(int64_t game_lo, int64_t game_hi, uint64_t seed, int64_t* stats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QTTT_B200_H */
