/*
 * tetris_b200.h -- C ABI of the B200-native batched Tetris environment.
 *
 * The reference (s0phia-/tetris) has no FFI layer: its boundary is the Python surface of
 * game.py / state.py / tetromino.py.  The Python package in this repo (tetris_b200.game.Tetris,
 * tetris_b200.batched.BatchedTetris, re-exported as package `tetris`) mirrors that surface and
 * binds the entry points below through ctypes; each entry point names the reference code it
 * replaces (file:line into the reference checkout).
 *
 * Conventions:
 *   - extern "C", plain pointers and sizes only; no C++ or torch types.
 *   - every call returns 0 on success, <0 on error; tb_last_error() gives the message
 *     (thread-local).  Nothing throws across the ABI.
 *   - all buffers are caller-owned DEVICE memory unless a parameter says "host"; nothing is
 *     allocated inside; launches are asynchronous on the caller's cudaStream_t (`stream`).
 *   - no global mutable state besides per-device __constant__ tables: safe from several host
 *     threads on distinct streams / devices.
 *   - board shape (num_columns C, num_rows R) selects a compiled template instantiation (one translation unit per
 *     shape, csrc/tb_shape.cu); tb_supported_shape() says which are loaded (built in: 10x20, 10x10, 6x12, 8x16, 4x4,
 *     16x27, 12x24).  The reference takes any num_columns / num_rows (game.py:21-31): any other shape with
 *     4 <= C <= 16 and 4 <= R <= 27 (uint16 row masks; 32-bit column masks with one spare bit) is compiled into its
 *     own shared object and added with tb_load_shape().
 *
 * Device state ("state" below) is one caller-owned allocation of tb_state_bytes() bytes, 256-byte
 * aligned, laid out as a structure of arrays over envs:
 *   planes  uint4[NB][n_env]   row masks, uint16 per row (bit c = cell (r, c)), 8 rows per 128-bit word,
 *                              NB = ceil((R + 4) / 8)
 *   meta    uint4[n_env]       bytes 0..9 column heights (lowest_free_rows, state.py:162-172; boards of up to 10
 *                              columns -- wider boards leave them zero, tb_export_boards derives heights from the
 *                              planes), byte 10 current piece (global id), byte 11 bag mask, bytes 12..15 draw counter
 *   epi     uint2[n_env]       placements and lines of the running episode
 *
 * Piece ids (global): 0 Straight, 1 RCorner, 2 LCorner, 3 Square, 4 SnakeR, 5 SnakeL, 6 T   (game.py:41-47)
 *                     7 ThreeL, 8 ThreeLine                                               (game.py:38-39)
 * Piece sets: 0 = reference default {ThreeL, ThreeLine}; 1 = the seven tetrominoes.
 * Afterstate slots are in the reference's enumeration (= action) order, tetromino.py:*.get_after_states.
 * Feature order (game.py:10-18): rows_with_holes, column_transitions, holes, landing_height,
 *                                cumulative_wells, row_transitions, eroded, hole_depth.
 */
#ifndef TETRIS_B200_H
#define TETRIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_VERSION 200          /* 0.2.0 */
#define TB_NUM_FEATURES 8
#define TB_MAX_SLOTS 60         /* ThreeL at C = 16 (36 at C = 10); per-slot masks are 64 bits */

/* tb_step / tb_afterstates flags */
#define TB_FLAG_AUTO_RESET     1   /* step: reset a finished env in place (what example_play.py:20-21 does) */
#define TB_FLAG_ACTION_IS_SLOT 2   /* step: actions index enumeration slots instead of non-terminal ranks   */
#define TB_FLAG_INCLUDE_TERMINAL 4 /* afterstates: also write the feature rows of terminal afterstates
                                      (get_after_states(include_terminal=True), game.py:74-78)               */
#define TB_FLAG_VALIDATE_ONLY  8   /* step: dry run -- only status_out is produced, no env is touched; lets a caller
                                      raise the reference's IndexError (game.py:83) BEFORE any env is stepped  */
#define TB_FLAG_FEATS_I16      16  /* afterstates: feats_out is int16[n_env][a_stride][8] holding 2 x feature
                                      (x direction): every feature is a half-integer below 2^14, so this is exact
                                      at half the bytes -- for policies on the host (PCIe) or bandwidth-bound readers */

/* tb_rollout policies */
#define TB_POLICY_RANDOM 0         /* uniformly random legal placement (per-env counter RNG, stream 1)       */
#define TB_POLICY_GREEDY 1         /* first arg-max of the float32 linear score over legal afterstates
                                      (Tetris.fitness, game.py:109-120)                                      */

/* tb_rollout statistics, int64 each; sums except the two maxima */
enum {
    TB_ST_PLACEMENTS = 0, TB_ST_EPISODES, TB_ST_LINES, TB_ST_REWARD, TB_ST_AFTERSTATES,
    TB_ST_LINES0, TB_ST_LINES1, TB_ST_LINES2, TB_ST_LINES3, TB_ST_LINES4,
    TB_ST_MAX_EP_LINES, TB_ST_MAX_EP_STEPS, TB_ST_SUM_EP_STEPS, TB_ST_SUM_EP_LINES,
    TB_ST_RESERVED0, TB_ST_RESERVED1, TB_ST_COUNT
};

int tb_version(void);
const char *tb_last_error(void);

/* 1 if kernels for this board shape are loaded (built in, or added with tb_load_shape). */
int tb_supported_shape(int num_columns, int num_rows);
/*
 * Add the kernels of one more board shape: `path` is a shared object built from csrc/tb_shape.cu with
 * -DTB_C=<num_columns> -DTB_R=<num_rows> -DTB_SHAPE_PLUGIN (tetris_b200._lib.build_shape does it).  The object stays
 * loaded for the life of the process.  Tetris(num_columns, num_rows) of the reference takes any size (game.py:21-31).
 */
int tb_load_shape(const char *path);
/*
 * Tuning knobs for experiments and tests; the defaults are what ships.  Names: "k1_cfg" / "k3_cfg" (tile configuration
 * of tb_afterstates / tb_rollout, -1 = chosen by batch size), "small_groups", "max_ctas" (0 = no cap; a small cap makes
 * every CTA loop over several tiles).  Defaults may also come from the environment (TB_K1_CFG, TB_K3_CFG,
 * TB_SMALL_GROUPS, TB_MAX_CTAS), which is read once.  tb_get_tuning returns the current value.
 */
int tb_set_tuning(const char *name, int value);
int tb_get_tuning(const char *name);
/* Bytes of device state for n_env envs (0 if the shape is unsupported). */
size_t tb_state_bytes(int num_columns, int num_rows, int64_t n_env);
/* Number of afterstates (enumeration slots) of a piece: SURVEY.md Appendix A. */
int tb_num_slots(int piece, int num_columns);
/*
 * The piece x rotation table entry behind enumeration slot `slot` (tetromino.py:33-576).
 * Host-only.  out int32[17]: [0] anchor_col  [1] width  [2] n_cells  [3] len(changed_lines)
 *   [4] 2 * landing_height_bonus  [5..8] pieces_per_changed_row  [9..16] the cells as (dx, dy) pairs.
 */
int tb_slot_info(int piece, int num_columns, int slot, int32_t *out);
/* max over the set's pieces = row stride callers should use for per-slot outputs. */
int tb_a_max(int num_columns, int piece_set);

/*
 * Tetris.__init__ + reset (game.py:21-63): empty boards, fresh bags, zero counters, draw the first piece.
 *   env_offset  global id of env 0 of this shard (per-env RNG is keyed by seed and GLOBAL env id, so results do
 *               not depend on how envs are sharded over GPUs)
 *   piece_tape  nullable uint8[n_env]: global piece ids to use instead of the bag RNG (parity runs); an id that
 *               names no piece leaves the env inert (no afterstates, never stepped)
 *   reset_mask  nullable uint8[n_env]: if given, behaves like Tetris.reset() (game.py:53-63) on the flagged
 *               envs only -- board emptied, one more piece drawn, the bag persists
 */
int tb_reset(void *state, int num_columns, int num_rows, int64_t n_env, int64_t env_offset, uint64_t seed,
             int piece_set, const uint8_t *piece_tape, const uint8_t *reset_mask, void *stream);

/*
 * Tetris.get_after_states (game.py:67-80) for every env: enumerate every rotation x column placement of the
 * current piece (tetromino.py:*.get_after_states), drop, lock, clear (state.py:121-143), terminal test
 * (state.py:111-117) and the eight BCTS features (state.py:97-107,175-280).
 *   feats_out   float32[n_env][a_stride][8], by enumeration slot, 16-byte aligned.  Rows of non-terminal afterstates
 *               (the ones game.py:69 keeps) are always written; rows of terminal afterstates only with
 *               TB_FLAG_INCLUDE_TERMINAL (game.py:74-78); rows >= tb_num_slots(piece) never.
 *   a_stride    row stride of feats_out in afterstates, 1..64; normally tb_a_max().  Slots >= a_stride are not
 *               written (their legality is still reported in valid_out / count_out).
 *   valid_out   nullable uint64[n_env]: bit s set  <=>  slot s is a non-terminal afterstate (a legal action)
 *   count_out   nullable int32[n_env]: number of legal actions (len(self.afterstates), game.py:69)
 *   directions  nullable HOST float[8]: per-feature multipliers (feature_directions, state.py:49-50), applied in fp32
 *   flags       TB_FLAG_INCLUDE_TERMINAL, TB_FLAG_FEATS_I16 (feats_out then is int16[n_env][a_stride][8] = 2 x feature)
 */
int tb_afterstates(const void *state, int num_columns, int num_rows, int64_t n_env, void *feats_out,
                   uint64_t *valid_out, int32_t *count_out, int a_stride, const float *directions, int flags,
                   void *stream);

/*
 * Afterstates with their boards, for the single-env State objects of the compatibility layer
 * (tetromino.py builds a State per afterstate, state.py:5-38).  All outputs nullable.
 *   rows_out    uint16[n_env][a_stride][R+4]   board after clearing
 *   heights_out uint8 [n_env][a_stride][C]     lowest_free_rows after clearing
 *   info_out    int32 [n_env][a_stride][4]     anchor_row, cleared-row mask (absolute rows), terminal, anchor_col
 */
int tb_afterstates_export(const void *state, int num_columns, int num_rows, int64_t n_env, float *feats_out,
                          uint16_t *rows_out, uint8_t *heights_out, int32_t *info_out, int a_stride, void *stream);

/*
 * Tetris.step (game.py:82-92) + is_game_over (game.py:94-100) for every env.
 *   actions     int32[n_env]: index into the non-terminal afterstates (game.py:69,83), or an enumeration slot
 *               with TB_FLAG_ACTION_IS_SLOT
 *   piece_tape  nullable uint8[n_env]: next piece per env (global id) instead of the bag RNG
 *   obs_out     nullable float32[n_env][8]: features of the chosen afterstate (game.py:91)
 *   reward_out  nullable int32[n_env]: lines - 1, and -100 more when done (game.py:86-90)
 *   done_out    nullable uint8[n_env]; lines_out nullable int32[n_env]
 *   status_out  nullable int32[1], zeroed by the caller: left 0 if every action was in range, else
 *               0x7FFFFFFF - (lowest env index with an out-of-range action).  The reference raises IndexError
 *               (game.py:83); here such an env is left untouched and its outputs are defined: zero observation,
 *               reward and lines, done = the env has no legal placement at all.  With TB_FLAG_VALIDATE_ONLY nothing
 *               but status_out is written and no env is stepped.
 */
int tb_step(void *state, int num_columns, int num_rows, int64_t n_env, int64_t env_offset, uint64_t seed,
            int piece_set, const int32_t *actions, const uint8_t *piece_tape, float *obs_out, int32_t *reward_out,
            uint8_t *done_out, int32_t *lines_out, int32_t *status_out, int flags, void *stream);

/*
 * The example_play.py:11-21 loop fused on the device: n_steps placements per env with an in-kernel policy,
 * game-over detection and auto-reset; episode statistics are ADDED into stats (max for the two maxima).
 *   weights     HOST float[8] for TB_POLICY_GREEDY (game.py:111-118 gives the BCTS weights); ignored for random
 *   stats       int64[TB_ST_COUNT] device
 */
int tb_rollout(void *state, int num_columns, int num_rows, int64_t n_env, int64_t env_offset, uint64_t seed,
               int piece_set, int n_steps, int policy, const float *weights, int64_t *stats, void *stream);

/*
 * Tetris.perform_rollouts / single_rollout (game.py:129-160) for every env and every action at once.  For each legal
 * enumeration slot s of env e and each of n_forks forks: take the action, draw the next piece, then follow the
 * in-kernel policy for `length - 1` more placements; the fork's return is -1 if the game ended on the way
 * (game.py:134,143-145), else the sum of the follow-up rewards (lines - 1 each, game.py:86,141).
 *   child_state caller-owned scratch of tb_state_bytes(C, R, n_env * a_stride * n_forks) bytes; child
 *               d = (e * a_stride + s) * n_forks + f draws pieces from the stream (seed2, child_offset + d), starting
 *               from the parent's bag (the reference forks share one global sampler and are not reproducible)
 *   piece_tape  nullable uint8[n_env * a_stride * n_forks][length]: the pieces child d draws, in order (the one drawn
 *               by the action itself first), instead of its RNG stream -- what a recorded sampler of the reference
 *               supplies (parity runs against game.py:129-160)
 *   ret_sum     int32[n_env][a_stride]: sum of the forks' returns (mean = ret_sum / n_forks); 0 for illegal slots
 *   valid_out   nullable uint64[n_env]: legal slots
 *   stats       int64[TB_ST_COUNT] device: statistics of the follow-up placements are added (as tb_rollout)
 */
int tb_rollout_values(const void *state, int num_columns, int num_rows, int64_t n_env, int piece_set, void *child_state,
                      int a_stride, int n_forks, int length, int policy, const float *weights, uint64_t seed2,
                      int64_t child_offset, const uint8_t *piece_tape, int32_t *ret_sum, uint64_t *valid_out,
                      int64_t *stats, void *stream);

/*
 * State interchange (state.py:22-25,162-172; utils.py:179-191 needs the board on the host).
 * export: rows_out uint16[count][R+4], heights_out uint8[count][C], piece_out uint8[count] (all nullable)
 * import: rows_in  uint16[count][R+4] (heights are recomputed), piece_in nullable uint8[count].  A board with a cell at
 *         or above row R is a terminal state (state.py:111-117): its env is marked finished (piece 0xFF) -- it has no
 *         afterstates and no kernel steps it.  A piece id that names no piece makes the env inert (0xFE).
 */
int tb_export_boards(const void *state, int num_columns, int num_rows, int64_t n_env, int64_t first, int64_t count,
                     uint16_t *rows_out, uint8_t *heights_out, uint8_t *piece_out, void *stream);
int tb_import_boards(void *state, int num_columns, int num_rows, int64_t n_env, int64_t first, int64_t count,
                     const uint16_t *rows_in, const uint8_t *piece_in, void *stream);

/*
 * State.__init__ + get_features on caller-supplied boards (state.py:5-38,97-107): clear the full rows among
 * changed_lines, recompute heights, terminal test, features.
 *   rows_in     uint16[n][R+4]
 *   params      int32[n][4]: anchor_row (changed_lines[0]), len(changed_lines) (1..4),
 *               pieces_per_changed_row packed 4 bits each, 2 * landing_height_bonus
 *   rows_out uint16[n][R+4], heights_out uint8[n][C], info_out int32[n][4] = (n_cleared, cleared-row mask,
 *   terminal, 0), feats_out float32[n][8]      (all nullable)
 */
int tb_eval_states(int num_columns, int num_rows, int64_t n, const uint16_t *rows_in, const int32_t *params,
                   uint16_t *rows_out, uint8_t *heights_out, int32_t *info_out, float *feats_out, void *stream);

/*
 * Tetris.fitness (game.py:109-120) for n feature rows: float32 products summed left to right, no FMA.
 *   feats float32[n][8] (device, 16-byte aligned), weights HOST float[8], out float32[n] (device)
 */
int tb_fitness(int64_t n, const float *feats, const float *weights, float *out, void *stream);

/*
 * Softmax policy over every env's legal afterstates and the policy-gradient term of a chosen action
 * (utils.py:26-38: compute_action_probabilities, grad_of_log_action_probabilities), float64 like the reference.
 *   feats     float32[n_env][a_stride][8] as written by tb_afterstates (16-byte aligned), valid uint64[n_env]
 *   weights   HOST double[8]; utilities = feats . weights / temperature, shifted by their maximum before exp
 *   actions   nullable int32[n_env]: chosen action as an enumeration slot (needed for grad_out)
 *   probs_out nullable double[n_env][a_stride]: probability per slot, 0 for illegal slots
 *   grad_out  nullable double[n_env][8]: feats[action] - sum_s probs[s] * feats[s]
 */
int tb_action_probabilities(int64_t n_env, int a_stride, const float *feats, const uint64_t *valid,
                            const double *weights, double temperature, const int32_t *actions, double *probs_out,
                            double *grad_out, void *stream);

/*
 * Combine n_parts episode-statistics vectors (int64[n_parts][TB_ST_COUNT], device) into out int64[TB_ST_COUNT]:
 * sums, except the two maxima.  The local half of the multi-GPU reduction: one all-gather of the ranks' vectors,
 * then this (tetris_b200.distributed.reduce_stats).
 */
int tb_combine_stats(const int64_t *parts, int n_parts, int64_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TETRIS_B200_H */
