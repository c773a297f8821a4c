// tb_abi.cu -- the C ABI of include/tetris_b200.h: argument checks, board-shape dispatch, and the few kernels that do
// not depend on the board shape.  The shape-specific kernels live in tb_kernels.cuh, one translation unit per shape
// (tb_shape.cu); this file only sees their launcher tables (TbShapeVT, tb_shape.h).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/tetris_b200.h"
#include "tb_core.cuh"
#include "tb_shape.h"

namespace tb {

// Sum of the rollout returns of an (env, action)'s forks: -1 for a fork that ended (game.py:134,143-145), else the
// lines cleared minus the placements made after the first step (reward = lines - 1 per step, game.py:86,141).
__global__ void k_fork_returns(const uint4 *__restrict__ meta, const uint2 *__restrict__ epi, int64_t n_parent,
                               int a_stride, int n_forks, int32_t *__restrict__ ret_sum,
                               unsigned long long *__restrict__ valid)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_parent) return;
    unsigned long long vm = 0ull;
    for (int s = 0; s < a_stride; ++s) {
        int sum = 0;
        bool legal = false;
        for (int f = 0; f < n_forks; ++f) {
            const int64_t d = (e * a_stride + s) * n_forks + f;
            const int piece = (int)((meta[d].z >> 16) & 0xffu);
            if (piece == kPieceVoid) continue;
            legal = true;
            const uint2 ep = epi[d];
            sum += piece == kPieceDead ? -1 : (int)ep.y - (int)ep.x;
        }
        ret_sum[e * a_stride + s] = sum;
        if (legal && s < 64) vm |= 1ull << s;
    }
    if (valid) valid[e] = vm;
}


// Tetris.fitness (game.py:109-120) of n feature rows
__global__ void k_fitness(int64_t n, const float *__restrict__ feats, F8 wts, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = reinterpret_cast<const float4 *>(feats)[2 * i], b = reinterpret_cast<const float4 *>(feats)[2 * i + 1];
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    out[i] = fitness(f, wts.v);
}

// Softmax policy over the legal afterstates of every env (utils.py:26-31 compute_action_probabilities) and the
// gradient of the log-probability of the chosen action (utils.py:35-38), float64 like the reference's NumPy.
// One thread per env; utilities are recomputed instead of stored (A <= 36 rows of 8 floats, L1/L2 resident).
struct D8 { double v[8]; };
__device__ __forceinline__ double utility(const float *__restrict__ row, const D8 &w, double inv_t)
{
    const float4 a = reinterpret_cast<const float4 *>(row)[0], b = reinterpret_cast<const float4 *>(row)[1];
    double u = (double)a.x * w.v[0];
    u += (double)a.y * w.v[1]; u += (double)a.z * w.v[2]; u += (double)a.w * w.v[3];
    u += (double)b.x * w.v[4]; u += (double)b.y * w.v[5]; u += (double)b.z * w.v[6]; u += (double)b.w * w.v[7];
    return u * inv_t;
}
__global__ void k_action_probs(int64_t n, int a_stride, const float *__restrict__ feats,
                               const unsigned long long *__restrict__ valid, D8 w, double inv_t,
                               const int32_t *__restrict__ actions, double *__restrict__ probs, double *__restrict__ grad)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned long long vm = valid[e];
    const float *base = feats + (size_t)e * a_stride * 8;
    double umax = -1.0e300;
    for (unsigned long long m = vm; m; m &= m - 1) umax = fmax(umax, utility(base + 8 * (__ffsll((long long)m) - 1), w, inv_t));
    double z = 0.0;
    for (unsigned long long m = vm; m; m &= m - 1) z += exp(utility(base + 8 * (__ffsll((long long)m) - 1), w, inv_t) - umax);
    const double inv_z = vm ? 1.0 / z : 0.0;
    double mean[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < a_stride; ++s) {
        double p = 0.0;
        if (s < 64 && ((vm >> s) & 1ull)) {
            const float *row = base + 8 * s;
            p = exp(utility(row, w, inv_t) - umax) * inv_z;
            if (grad) {
#pragma unroll
                for (int i = 0; i < 8; ++i) mean[i] += p * (double)row[i];
            }
        }
        if (probs) probs[(size_t)e * a_stride + s] = p;
    }
    if (grad) {
        const int sel = actions ? actions[e] : -1;            // enumeration slot of the chosen action
        const bool ok = sel >= 0 && sel < a_stride && sel < 64 && ((vm >> sel) & 1ull);
#pragma unroll
        for (int i = 0; i < 8; ++i) grad[e * 8 + i] = ok ? (double)base[8 * sel + i] - mean[i] : 0.0;
    }
}


// Combine per-shard statistics vectors (tb_rollout's int64[TB_ST_COUNT]): sums, except the two maxima.  One warp.
__global__ void k_combine_stats(const long long *__restrict__ parts, int n_parts, long long *__restrict__ out)
{
    const int i = threadIdx.x;
    if (i >= TB_ST_COUNT) return;
    const bool is_max = (i == TB_ST_MAX_EP_LINES || i == TB_ST_MAX_EP_STEPS);
    long long acc = 0;
    for (int p = 0; p < n_parts; ++p) {
        const long long v = parts[(size_t)p * TB_ST_COUNT + i];
        acc = is_max ? (v > acc ? v : acc) : acc + v;
    }
    out[i] = acc;
}

}  // namespace tb

// =============================================================================================
// shape registry, tuning, errors
// =============================================================================================
using namespace tb;

// Board shapes linked into this library: the three of BASELINE.json's configs, two extras (mid-size, tiny edge case)
// and two beyond the reference's usual sizes (wide / tall: up to 16 columns and 27 rows fit the uint16 row masks and
// 32-bit column masks).  Any other shape with 4 <= C <= 16, 4 <= R <= 27 can be compiled into its own shared object
// (tetris_b200._lib.build_shape) and added with tb_load_shape().
// The list is generated by the build (tetris_b200/_lib.py: BUILTIN_SHAPES -> build/tb_builtin_shapes.inc).
#if !defined(TB_BUILTIN_SHAPES) && __has_include("build/tb_builtin_shapes.inc")
#include "build/tb_builtin_shapes.inc"
#endif
#ifndef TB_BUILTIN_SHAPES
#define TB_BUILTIN_SHAPES(X) X(10, 20) X(10, 10) X(6, 12) X(8, 16) X(4, 4) X(16, 27) X(12, 24)
#endif
#define X(c, r) extern "C" const TbShapeVT *tb_shape_vt_##c##x##r(void);
TB_BUILTIN_SHAPES(X)
#undef X

static constexpr int kMaxShapes = 64;
static const TbShapeVT *g_shapes[kMaxShapes];
static std::atomic<int> g_n_shapes{0};
static std::mutex g_shape_mu;

static int register_shape(const TbShapeVT *vt)
{
    if (!vt || vt->abi != TB_SHAPE_ABI) return -1;
    std::lock_guard<std::mutex> lock(g_shape_mu);
    const int n = g_n_shapes.load(std::memory_order_relaxed);
    for (int i = 0; i < n; ++i)
        if (g_shapes[i]->C == vt->C && g_shapes[i]->R == vt->R) return 0;      // first registration wins
    if (n >= kMaxShapes) return -1;
    g_shapes[n] = vt;
    g_n_shapes.store(n + 1, std::memory_order_release);
    return 0;
}
static void register_builtin_shapes()
{
    static std::once_flag once;
    std::call_once(once, [] {
#define X(c, r) register_shape(tb_shape_vt_##c##x##r());
        TB_BUILTIN_SHAPES(X)
#undef X
    });
}
static const TbShapeVT *find_shape(int C, int R)
{
    register_builtin_shapes();
    const int n = g_n_shapes.load(std::memory_order_acquire);
    for (int i = 0; i < n; ++i)
        if (g_shapes[i]->C == C && g_shapes[i]->R == R) return g_shapes[i];
    return nullptr;
}

static thread_local char g_err[256] = "";
static int fail(const char *fmt, const char *detail)
{
    snprintf(g_err, sizeof g_err, fmt, detail);
    return -1;
}
static int check_launch(const char *what)
{
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}
static int sm_count()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

// Tuning knobs (experiments and tests only; the defaults are what ships).  Read from the environment ONCE, when the
// first launch needs them (TB_K1_CFG, TB_K3_CFG, TB_SMALL_GROUPS, TB_MAX_CTAS), and settable through tb_set_tuning().
enum { TUNE_K1_CFG = 0, TUNE_K3_CFG, TUNE_SMALL_GROUPS, TUNE_MAX_CTAS, TUNE_K2_CFG, TUNE_K3R_CFG, TUNE_COUNT };
static const char *const kTuneNames[TUNE_COUNT] = {"k1_cfg", "k3_cfg", "small_groups", "max_ctas", "k2_cfg", "k3r_cfg"};
static const char *const kTuneEnv[TUNE_COUNT] = {"TB_K1_CFG", "TB_K3_CFG", "TB_SMALL_GROUPS", "TB_MAX_CTAS", "TB_K2_CFG", "TB_K3R_CFG"};
static const int kTuneDefault[TUNE_COUNT] = {-1, -1, 4, 0, -1, -1};
static std::atomic<int> g_tune[TUNE_COUNT];
static void init_tuning()
{
    static std::once_flag once;
    std::call_once(once, [] {
        for (int i = 0; i < TUNE_COUNT; ++i) {
            const char *v = getenv(kTuneEnv[i]);
            g_tune[i].store(v ? atoi(v) : kTuneDefault[i], std::memory_order_relaxed);
        }
    });
}
static TbLaunchCtx make_ctx(void *stream)
{
    init_tuning();
    TbLaunchCtx cx;
    cx.stream = stream;
    cx.sm_count = sm_count();
    cx.k1_cfg = g_tune[TUNE_K1_CFG].load(std::memory_order_relaxed);
    cx.k3_cfg = g_tune[TUNE_K3_CFG].load(std::memory_order_relaxed);
    cx.k2_cfg = g_tune[TUNE_K2_CFG].load(std::memory_order_relaxed);
    cx.k3r_cfg = g_tune[TUNE_K3R_CFG].load(std::memory_order_relaxed);
    cx.small_groups = g_tune[TUNE_SMALL_GROUPS].load(std::memory_order_relaxed);
    cx.max_ctas = g_tune[TUNE_MAX_CTAS].load(std::memory_order_relaxed);
    cx.err = g_err;
    cx.err_len = sizeof g_err;
    return cx;
}

extern "C" {

int tb_version(void) { return TB_VERSION; }
const char *tb_last_error(void) { return g_err; }

int tb_supported_shape(int C, int R) { return find_shape(C, R) != nullptr; }

int tb_load_shape(const char *path)
{
    if (!path) return fail("%s: null path", __func__);
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return fail("tb_load_shape: %s", dlerror());
    TbShapeGetter get = (TbShapeGetter)dlsym(h, "tb_shape_vt");
    if (!get) { dlclose(h); return fail("%s: the object does not export tb_shape_vt", __func__); }
    const TbShapeVT *vt = get();
    if (!vt || vt->abi != TB_SHAPE_ABI) { dlclose(h); return fail("%s: shape object built for another version", __func__); }
    register_builtin_shapes();
    if (register_shape(vt)) { dlclose(h); return fail("%s: shape table is full", __func__); }
    return 0;                                              // the object stays loaded for the life of the process
}

int tb_set_tuning(const char *name, int value)
{
    init_tuning();
    for (int i = 0; name && i < TUNE_COUNT; ++i)
        if (!strcmp(name, kTuneNames[i])) { g_tune[i].store(value, std::memory_order_relaxed); return 0; }
    return fail("%s: unknown tuning name", __func__);
}
int tb_get_tuning(const char *name)
{
    init_tuning();
    for (int i = 0; name && i < TUNE_COUNT; ++i)
        if (!strcmp(name, kTuneNames[i])) return g_tune[i].load(std::memory_order_relaxed);
    return -0x7FFFFFFF;
}

size_t tb_state_bytes(int C, int R, int64_t n_env)
{
    const TbShapeVT *vt = find_shape(C, R);
    return vt ? (size_t)n_env * vt->bytes_per_env : 0;
}

int tb_num_slots(int piece, int C)
{
    if (piece < 0 || piece >= kNumPieces) return 0;
    return piece_num_slots(kPieceHost[piece], C);
}

int tb_a_max(int C, int piece_set)
{
    int m = 0;
    for (int i = 0; i < set_size(piece_set); ++i) {
        const int n = tb_num_slots(set_piece(piece_set, i), C);
        if (n > m) m = n;
    }
    return m;
}

int tb_slot_info(int piece, int C, int slot, int32_t *out)
{
    if (piece < 0 || piece >= kNumPieces || !out) return fail("%s: bad piece or null output", __func__);
    const uint32_t pw = kPieceHost[piece];
    if (C < 4 || slot < 0 || slot >= piece_num_slots(pw, C)) return fail("%s: slot out of range", __func__);
    int ori, c;
    slot_to_placement(pw, C, slot, ori, c);
    const uint32_t d = kOriHost[ori];
    int n = 0;
    for (int i = 0; i < 17; ++i) out[i] = 0;
    out[0] = c; out[1] = desc_w(d); out[3] = desc_chg(d); out[4] = desc_bonus2(d);
    for (int dx = 0; dx < 4; ++dx)
        for (int k = 0; k < desc_len(d, dx); ++k) {
            const int dy = desc_bot(d, dx) + k;
            out[9 + 2 * n] = dx; out[10 + 2 * n] = dy;
            if (dy < desc_chg(d)) out[5 + dy] += 1;                 // pieces_per_changed_row
            ++n;
        }
    out[2] = n;
    return 0;
}

#define TB_CHECK_COMMON()                                                                   \
    if (n_env <= 0) return fail("%s: n_env must be positive", __func__);                   \
    const TbShapeVT *vt = find_shape(C, R);                                                 \
    if (!vt) return fail("%s: unsupported board shape (see tb_supported_shape / tb_load_shape)", __func__);

int tb_reset(void *state, int C, int R, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
             const uint8_t *piece_tape, const uint8_t *reset_mask, void *stream)
{
    TB_CHECK_COMMON();
    if (piece_set < 0 || piece_set > 1) return fail("%s: piece_set must be 0 or 1", __func__);
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->reset(&cx, state, n_env, env_offset, seed, piece_set, piece_tape, reset_mask);
}

int tb_afterstates(const void *state, int C, int R, int64_t n_env, void *feats_out, uint64_t *valid_out,
                   int32_t *count_out, int a_stride, const float *directions, int flags, void *stream)
{
    TB_CHECK_COMMON();
    if (!feats_out) return fail("%s: feats_out is required", __func__);
    if (a_stride < 1 || a_stride > 64) return fail("%s: a_stride must be in 1..64 (slots >= a_stride are not written)", __func__);
    if ((reinterpret_cast<uintptr_t>(feats_out) & 15u) != 0) return fail("%s: feats_out must be 16-byte aligned", __func__);
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->afterstates(&cx, state, n_env, feats_out, valid_out, count_out, a_stride, directions, flags);
}

int tb_afterstates_export(const void *state, int C, int R, int64_t n_env, float *feats_out, uint16_t *rows_out,
                          uint8_t *heights_out, int32_t *info_out, int a_stride, void *stream)
{
    TB_CHECK_COMMON();
    if (a_stride < 1 || a_stride > 64) return fail("%s: a_stride must be in 1..64 (slots >= a_stride are not written)", __func__);
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->afterstates_export(&cx, state, n_env, feats_out, rows_out, heights_out, info_out, a_stride);
}

int tb_step(void *state, int C, int R, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
            const int32_t *actions, const uint8_t *piece_tape, float *obs_out, int32_t *reward_out, uint8_t *done_out,
            int32_t *lines_out, int32_t *status_out, int flags, void *stream)
{
    TB_CHECK_COMMON();
    if (!actions) return fail("%s: actions is required", __func__);
    if (piece_set < 0 || piece_set > 1) return fail("%s: piece_set must be 0 or 1", __func__);
    if ((flags & TB_FLAG_VALIDATE_ONLY) && !status_out) return fail("%s: TB_FLAG_VALIDATE_ONLY needs status_out", __func__);
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->step(&cx, state, n_env, env_offset, seed, piece_set, actions, piece_tape, obs_out, reward_out, done_out,
                    lines_out, status_out, flags);
}

static int check_rollout_args(const char *fn, int piece_set, int n_steps, int policy, const float *weights, const int64_t *stats)
{
    if (!stats) return fail("%s: stats is required", fn);
    if (n_steps < 0) return fail("%s: n_steps must be >= 0", fn);
    if (piece_set < 0 || piece_set > 1) return fail("%s: piece_set must be 0 or 1", fn);
    if (policy == TB_POLICY_GREEDY && !weights) return fail("%s: greedy policy needs weights", fn);
    if (policy != TB_POLICY_GREEDY && policy != TB_POLICY_RANDOM) return fail("%s: unknown policy", fn);
    return 0;
}

int tb_rollout(void *state, int C, int R, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
               int n_steps, int policy, const float *weights, int64_t *stats, void *stream)
{
    TB_CHECK_COMMON();
    if (check_rollout_args(__func__, piece_set, n_steps, policy, weights, stats)) return -1;
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->rollout(&cx, state, n_env, env_offset, seed, piece_set, n_steps, policy, weights, stats, 0, nullptr, 0);
}

int tb_rollout_values(const void *state, int C, int R, int64_t n_env, int piece_set, void *child_state, int a_stride,
                      int n_forks, int length, int policy, const float *weights, uint64_t seed2, int64_t child_offset,
                      const uint8_t *piece_tape, int32_t *ret_sum, uint64_t *valid_out, int64_t *stats, void *stream)
{
    TB_CHECK_COMMON();
    if (!child_state || !ret_sum || !stats) return fail("%s: child_state, ret_sum and stats are required", __func__);
    if (a_stride < 1 || a_stride > 64 || n_forks < 1 || length < 1) return fail("%s: bad a_stride / n_forks / length", __func__);
    if (check_rollout_args(__func__, piece_set, length - 1, policy, weights, stats)) return -1;
    const int64_t n_child = n_env * a_stride * n_forks;
    const TbLaunchCtx cx = make_ctx(stream);
    int rc = vt->fork(&cx, state, n_env, child_state, a_stride, n_forks, seed2, child_offset, piece_set, piece_tape, length);
    if (rc) return rc;
    if (length > 1) {
        rc = vt->rollout(&cx, child_state, n_child, child_offset, seed2, piece_set, length - 1, policy, weights, stats, 1,
                         piece_tape ? piece_tape + 1 : nullptr, length);
        if (rc) return rc;
    }
    // child state layout (include/tetris_b200.h): planes uint4[NB][n], meta uint4[n], epi uint2[n]
    const size_t plane_bytes = vt->bytes_per_env - 24;
    const char *base = (const char *)child_state;
    const uint4 *meta = (const uint4 *)(base + plane_bytes * (size_t)n_child);
    const uint2 *epi = (const uint2 *)(base + (plane_bytes + 16) * (size_t)n_child);
    k_fork_returns<<<(unsigned)((n_env + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        meta, epi, n_env, a_stride, n_forks, ret_sum, (unsigned long long *)valid_out);
    return check_launch("tb_rollout_values(reduce)");
}

int tb_export_boards(const void *state, int C, int R, int64_t n_env, int64_t first, int64_t count,
                     uint16_t *rows_out, uint8_t *heights_out, uint8_t *piece_out, void *stream)
{
    TB_CHECK_COMMON();
    if (first < 0 || count < 0 || first + count > n_env) return fail("%s: env range out of bounds", __func__);
    if (count == 0) return 0;
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->export_boards(&cx, state, n_env, first, count, rows_out, heights_out, piece_out);
}

int tb_import_boards(void *state, int C, int R, int64_t n_env, int64_t first, int64_t count,
                     const uint16_t *rows_in, const uint8_t *piece_in, void *stream)
{
    TB_CHECK_COMMON();
    if (first < 0 || count < 0 || first + count > n_env) return fail("%s: env range out of bounds", __func__);
    if (!rows_in) return fail("%s: rows_in is required", __func__);
    if (count == 0) return 0;
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->import_boards(&cx, state, n_env, first, count, rows_in, piece_in);
}

int tb_eval_states(int C, int R, int64_t n, const uint16_t *rows_in, const int32_t *params, uint16_t *rows_out,
                   uint8_t *heights_out, int32_t *info_out, float *feats_out, void *stream)
{
    if (n <= 0) return fail("%s: n must be positive", __func__);
    const TbShapeVT *vt = find_shape(C, R);
    if (!vt) return fail("%s: unsupported board shape (see tb_supported_shape / tb_load_shape)", __func__);
    if (!rows_in) return fail("%s: rows_in is required", __func__);
    const TbLaunchCtx cx = make_ctx(stream);
    return vt->eval_states(&cx, n, rows_in, params, rows_out, heights_out, info_out, feats_out);
}

int tb_fitness(int64_t n, const float *feats, const float *weights, float *out, void *stream)
{
    if (n <= 0) return fail("%s: n must be positive", __func__);
    if (!feats || !weights || !out) return fail("%s: null argument", __func__);
    if ((reinterpret_cast<uintptr_t>(feats) & 15u) != 0) return fail("%s: feats must be 16-byte aligned", __func__);
    F8 w;
    for (int i = 0; i < 8; ++i) w.v[i] = weights[i];
    k_fitness<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, feats, w, out);
    return check_launch("tb_fitness");
}

int tb_action_probabilities(int64_t n_env, int a_stride, const float *feats, const uint64_t *valid,
                            const double *weights, double temperature, const int32_t *actions, double *probs_out,
                            double *grad_out, void *stream)
{
    if (n_env <= 0) return fail("%s: n_env must be positive", __func__);
    if (!feats || !valid || !weights) return fail("%s: feats, valid and weights are required", __func__);
    if (a_stride < 1 || a_stride > 64) return fail("%s: a_stride must be in 1..64", __func__);
    if (!(temperature > 0.0)) return fail("%s: temperature must be positive", __func__);
    if ((reinterpret_cast<uintptr_t>(feats) & 15u) != 0) return fail("%s: feats must be 16-byte aligned", __func__);
    D8 w;
    for (int i = 0; i < 8; ++i) w.v[i] = weights[i];
    k_action_probs<<<(unsigned)((n_env + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        n_env, a_stride, feats, (const unsigned long long *)valid, w, 1.0 / temperature, actions, probs_out, grad_out);
    return check_launch("tb_action_probabilities");
}

int tb_combine_stats(const int64_t *parts, int n_parts, int64_t *out, void *stream)
{
    if (!parts || !out || n_parts < 1) return fail("%s: parts, out and n_parts >= 1 are required", __func__);
    k_combine_stats<<<1, 32, 0, (cudaStream_t)stream>>>((const long long *)parts, n_parts, (long long *)out);
    return check_launch("tb_combine_stats");
}

}  // extern "C"
