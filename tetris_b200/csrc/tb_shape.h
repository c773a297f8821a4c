// tb_shape.h -- the table of launchers one board shape <C, R> exports (internal; the public ABI is include/tetris_b200.h).
//
// Every kernel is a template over the board shape, so each shape is its own translation unit (tb_shape.cu compiled with
// -DTB_C=.. -DTB_R=..) exporting `const TbShapeVT *tb_shape_vt_<C>x<R>(void)`.  tb_abi.cu registers the shapes linked
// into libtetris_b200.so and, through tb_load_shape(), shapes compiled later into their own shared objects
// (the reference takes any num_columns / num_rows, game.py:21-31).  Plain C types only: a shape object has no link-time
// dependency on the ABI object.
#ifndef TB_SHAPE_H
#define TB_SHAPE_H

#include <stddef.h>
#include <stdint.h>

#define TB_SHAPE_ABI 5          /* bumped whenever TbShapeVT / TbLaunchCtx change */

/* what every launcher needs besides its arguments */
typedef struct TbLaunchCtx {
    void *stream;               /* cudaStream_t */
    int sm_count;
    int k1_cfg, k3_cfg;         /* tile configuration of K1 / K3: -1 = by batch size (what ships), else forced (tests) */
    int k2_cfg, k3r_cfg;        /* CTA shape of K2 / of the random rollout: -1 / 0 = default */
    int small_groups;           /* 32-env groups per SM up to which the small-batch configuration is used */
    int max_ctas;               /* 0 = no cap; tests cap the grid so that every CTA loops over several tiles */
    char *err;                  /* error message buffer (thread-local in the ABI object) */
    size_t err_len;
} TbLaunchCtx;

typedef struct TbShapeVT {
    int abi;                    /* TB_SHAPE_ABI */
    int C, R;
    size_t bytes_per_env;       /* 16 * (NB + 1) + 8 */
    int (*reset)(const TbLaunchCtx *, void *state, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
                 const uint8_t *tape, const uint8_t *mask);
    int (*afterstates)(const TbLaunchCtx *, const void *state, int64_t n_env, void *feats_out, uint64_t *valid_out,
                       int32_t *count_out, int a_stride, const float *directions, int flags);
    int (*afterstates_export)(const TbLaunchCtx *, const void *state, int64_t n_env, float *feats_out, uint16_t *rows_out,
                              uint8_t *heights_out, int32_t *info_out, int a_stride);
    int (*step)(const TbLaunchCtx *, void *state, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
                const int32_t *actions, const uint8_t *tape, float *obs, int32_t *reward, uint8_t *done, int32_t *lines,
                int32_t *status, int flags);
    /* tape (nullable): uint8[n_env][tape_stride], the piece drawn by placement t of this launch is tape[e][t];
       only with no_reset (the rollout forks of tb_rollout_values) */
    int (*rollout)(const TbLaunchCtx *, void *state, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
                   int n_steps, int policy, const float *weights, int64_t *stats, int no_reset, const uint8_t *tape,
                   int tape_stride);
    int (*fork)(const TbLaunchCtx *, const void *parent, int64_t n_env, void *child, int a_stride, int n_forks,
                uint64_t seed2, int64_t child_offset, int piece_set, const uint8_t *tape, int tape_stride);
    int (*export_boards)(const TbLaunchCtx *, const void *state, int64_t n_env, int64_t first, int64_t count,
                         uint16_t *rows_out, uint8_t *heights_out, uint8_t *piece_out);
    int (*import_boards)(const TbLaunchCtx *, void *state, int64_t n_env, int64_t first, int64_t count,
                         const uint16_t *rows_in, const uint8_t *piece_in);
    int (*eval_states)(const TbLaunchCtx *, int64_t n, const uint16_t *rows_in, const int32_t *params, uint16_t *rows_out,
                       uint8_t *heights_out, int32_t *info_out, float *feats_out);
} TbShapeVT;

typedef const TbShapeVT *(*TbShapeGetter)(void);

#endif /* TB_SHAPE_H */
