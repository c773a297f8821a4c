// tb_kernels.cu -- sm_100a kernels + C ABI (include/tetris_b200.h) of the batched Tetris environment.
//
// Kernels:
//   k_reset        K0  Tetris.__init__/reset            game.py:21-63
//   k_afterstates  K1  Tetris.get_after_states           game.py:67-80 -> tetromino.py -> state.py
//   k_step         K2  Tetris.step + is_game_over        game.py:82-100
//   k_rollout_*    K3  example_play.py:11-21 loop fused with an in-kernel policy
//
// Mapping.  The work is integer/bit manipulation on a few dozen bytes per env -- no GEMM shape anywhere, so
// no tensor cores.  Envs are independent; a warp owns a tile of 32 envs.  Per-env work (load + transpose,
// the env record, applying a placement, RNG, game-over) runs one lane per env; per-afterstate work runs one
// lane per (env, slot) item over the tile's flattened item list (~23 items per env on average), so lanes
// stay busy although pieces have 9..34 placements.  Placements that clear a line or reach the top are
// rare and costly (from-scratch evaluation): they are queued in shared memory and evaluated 32 at a time
// instead of diverging the common incremental path.
//
// HBM layout: see include/tetris_b200.h (row masks, 8 rows per 128-bit word, SoA over envs -> every global
// load/store of state is a coalesced 128-bit access).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tetris_b200.h"
#include "tb_core.cuh"

namespace tb {

__constant__ uint32_t c_ori[kNumOris] = { TB_ORI_TABLE(TB_X_ORI) };
__constant__ uint32_t c_piece[kNumPieces] = { TB_PIECE_TABLE(TB_X_PIECE) };

constexpr unsigned FULLMASK = 0xFFFFFFFFu;

struct StateView {
    uint4 *planes;      // [NB][n_env]
    uint4 *meta;        // [n_env]
    uint2 *epi;         // [n_env]
    int64_t n_env;
};
struct F8 { float v[8]; };

template <int C, int R>
static StateView make_view(const void *base, int64_t n_env)
{
    using S = Shape<C, R>;
    char *p = (char *)base;
    StateView v;
    v.planes = (uint4 *)p;
    v.meta = (uint4 *)(p + (size_t)16 * S::NB * (size_t)n_env);
    v.epi = (uint2 *)(p + (size_t)16 * (S::NB + 1) * (size_t)n_env);
    v.n_env = n_env;
    return v;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
struct Meta { int piece; uint32_t bag, draws; };

__device__ __forceinline__ Meta unpack_meta(uint4 m)
{
    Meta r;
    r.piece = (int)((m.z >> 16) & 0xffu);
    r.bag = (m.z >> 24) & 0xffu;
    r.draws = m.w;
    return r;
}
template <int C>
__device__ __forceinline__ uint4 pack_meta(const uint32_t *col, Meta mt)
{
    uint32_t w[3] = {0u, 0u, 0u};
#pragma unroll
    for (int c = 0; c < C; ++c) w[c >> 2] |= (uint32_t)height_of(col[c]) << (8 * (c & 3));
    return make_uint4(w[0], w[1], w[2] | ((uint32_t)mt.piece << 16) | (mt.bag << 24), mt.draws);
}
template <int C, int R>
__device__ __forceinline__ void load_board(const StateView &sv, int64_t e, uint32_t *col)
{
    using S = Shape<C, R>;
    uint32_t w[S::NW];
#pragma unroll
    for (int b = 0; b < S::NB; ++b) {
        const uint4 v = sv.planes[(int64_t)b * sv.n_env + e];
        w[4 * b] = v.x; w[4 * b + 1] = v.y; w[4 * b + 2] = v.z; w[4 * b + 3] = v.w;
    }
    rows_to_cols<C, R>(w, col);
}
template <int C, int R>
__device__ __forceinline__ void store_board(const StateView &sv, int64_t e, const uint32_t *col)
{
    using S = Shape<C, R>;
    uint32_t w[S::NW];
    cols_to_rows<C, R>(col, w);
#pragma unroll
    for (int b = 0; b < S::NB; ++b)
        sv.planes[(int64_t)b * sv.n_env + e] = make_uint4(w[4 * b], w[4 * b + 1], w[4 * b + 2], w[4 * b + 3]);
}
__device__ __forceinline__ void stage_tables(uint32_t *s_ori, uint32_t *s_piece)
{
    if (threadIdx.x < kNumOris) s_ori[threadIdx.x] = c_ori[threadIdx.x];
    if (threadIdx.x < kNumPieces) s_piece[threadIdx.x] = c_piece[threadIdx.x];
    __syncthreads();
}
// draw the next piece of env (tape, or the env's bag RNG); returns the global piece id
__device__ __forceinline__ int draw_piece(int piece_set, uint64_t key, Meta &mt, const uint8_t *tape, int64_t e)
{
    if (tape) { mt.draws += 1; return (int)tape[e]; }
    return set_piece(piece_set, bag_draw(set_size(piece_set), key, mt.bag, mt.draws));
}
template <int C>
__device__ __forceinline__ int max_height(const uint32_t *col)
{
    uint32_t any = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) any |= col[c];
    return height_of(any);
}
// does the piece have at least one legal placement on this board?  (is_game_over, game.py:94-100)
template <int C, int R>
__device__ __forceinline__ bool any_valid(const uint32_t *col, uint32_t pw, const uint32_t *s_ori)
{
    if (max_height<C>(col) + 4 <= R) return true;       // every piece is at most 4 rows tall
    const int n = piece_num_slots(pw, C);
    for (int s = 0; s < n; ++s) {
        int ori, c;
        slot_to_placement(pw, C, s, ori, c);
        if (placement_valid<C, R>(col, s_ori[ori], c)) return true;
    }
    return false;
}
// mask of legal slots
template <int C, int R>
__device__ __forceinline__ unsigned long long valid_mask(const uint32_t *col, uint32_t pw, const uint32_t *s_ori)
{
    const int n = piece_num_slots(pw, C);
    if (max_height<C>(col) + 4 <= R) return (1ull << n) - 1ull;
    unsigned long long m = 0;
    for (int s = 0; s < n; ++s) {
        int ori, c;
        slot_to_placement(pw, C, s, ori, c);
        if (placement_valid<C, R>(col, s_ori[ori], c)) m |= 1ull << s;
    }
    return m;
}
__device__ __forceinline__ int nth_set_bit(unsigned long long m, int n)
{
    for (int i = 0; i < n; ++i) m &= m - 1ull;
    return __ffsll((long long)m) - 1;
}

// ---------------------------------------------------------------------------------------------
// K0 reset
// ---------------------------------------------------------------------------------------------
template <int C, int R>
__global__ void __launch_bounds__(256)
k_reset(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, const uint8_t *__restrict__ tape,
        const uint8_t *__restrict__ mask)
{
    using S = Shape<C, R>;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= sv.n_env) return;
    Meta mt;
    if (mask) {
        if (!mask[e]) return;
        mt = unpack_meta(sv.meta[e]);
    } else {
        mt.bag = 0u; mt.draws = 0u;
    }
#pragma unroll
    for (int b = 0; b < S::NB; ++b) sv.planes[(int64_t)b * sv.n_env + e] = make_uint4(0u, 0u, 0u, 0u);
    mt.piece = draw_piece(piece_set, env_key(seed, (uint64_t)(env_offset + e)), mt, tape, e);
    uint32_t zero[C];
#pragma unroll
    for (int c = 0; c < C; ++c) zero[c] = 0u;
    sv.meta[e] = pack_meta<C>(zero, mt);
    sv.epi[e] = make_uint2(0u, 0u);
}

// ---------------------------------------------------------------------------------------------
// K1 afterstates
// ---------------------------------------------------------------------------------------------
template <int C, int R>
struct TileSmem {
    using K = Rec<C, R>;
    uint32_t rec[32 * K::WORDS];
    unsigned long long acc[32];      // K1: legal-slot bits set by the slow path; K3: best (score, slot) key per env
    uint16_t pref[34];
    uint16_t queue[64];
    uint8_t pid[32];
};

// piece/orientation tables and the run-sum table into shared memory (one barrier)
template <int R>
__device__ __forceinline__ void stage_all(uint32_t *s_ori, uint32_t *s_piece, uint16_t *s_run)
{
    if (threadIdx.x < kNumOris) s_ori[threadIdx.x] = c_ori[threadIdx.x];
    if (threadIdx.x < kNumPieces) s_piece[threadIdx.x] = c_piece[threadIdx.x];
    for (int m = threadIdx.x; m < RunTab<R>::SIZE; m += blockDim.x) s_run[m] = run_tab_entry<R>((uint32_t)m);
    __syncthreads();
}

__device__ __forceinline__ void emit_features(float *__restrict__ feats, int64_t env, int a_stride, int slot,
                                              const Eval &ev, const F8 &dirs)
{
    float4 *dst = reinterpret_cast<float4 *>(feats + ((size_t)env * (size_t)a_stride + (size_t)slot) * 8);
    dst[0] = make_float4(ev.f[0] * dirs.v[0], ev.f[1] * dirs.v[1], ev.f[2] * dirs.v[2], ev.f[3] * dirs.v[3]);
    dst[1] = make_float4(ev.f[4] * dirs.v[4], ev.f[5] * dirs.v[5], ev.f[6] * dirs.v[6], ev.f[7] * dirs.v[7]);
}

// Flattened item list of a tile: item i belongs to the env whose slot range [pref[env], pref[env+1]) holds i.
// `heads` has bit j set when item base+j is the first item of an env; `cum` = envs that start before `base`.
__device__ __forceinline__ uint32_t window_heads(int own_lo, int own_n, int base)
{
    const unsigned rel = (unsigned)(own_lo - base);
    return __reduce_or_sync(FULLMASK, (own_n > 0 && rel < 32u) ? (1u << rel) : 0u);
}
// bits of a per-item ballot that fall into the owner lane's slot range, moved to slot positions
__device__ __forceinline__ unsigned long long window_bits(uint32_t bal, int own_lo, int own_hi, int base)
{
    const int s = imax(own_lo, base), t = imin(own_hi, base + 32);
    if (s >= t) return 0ull;
    const uint32_t bits = (bal >> (s - base)) & (0xFFFFFFFFu >> (32 - (t - s)));
    return (unsigned long long)bits << (s - own_lo);
}

template <int C, int R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_afterstates(StateView sv, float *__restrict__ feats, unsigned long long *__restrict__ valid_out,
              int *__restrict__ count_out, int a_stride, F8 dirs, int flags)
{
    using K = Rec<C, R>;
    __shared__ TileSmem<C, R> s_tile[WARPS];
    __shared__ uint32_t s_ori[32], s_piece[16];
    __shared__ uint16_t s_run[RunTab<R>::SIZE];
    stage_all<R>(s_ori, s_piece, s_run);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lane_le = (2u << lane) - 1u;
    TileSmem<C, R> &sm = s_tile[warp];
    uint32_t *acc32 = reinterpret_cast<uint32_t *>(sm.acc);
    const int64_t n_tiles = (sv.n_env + 31) >> 5;
    const bool want_terminal = (flags & TB_FLAG_INCLUDE_TERMINAL) != 0;

    for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * WARPS) {
        const int64_t e0 = tile * 32, e = e0 + lane;
        // ---- phase A: one lane per env: load, transpose, build the env record
        int n_slots = 0;
        if (e < sv.n_env) {
            uint32_t col[C];
            load_board<C, R>(sv, e, col);
            const Meta mt = unpack_meta(sv.meta[e]);
            build_env<C, R>(col, sm.rec + lane * K::WORDS);
            sm.pid[lane] = (uint8_t)mt.piece;
            n_slots = piece_num_slots(s_piece[mt.piece], C);
        }
        sm.acc[lane] = 0ull;
        int incl = n_slots;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULLMASK, incl, o);
            if (lane >= o) incl += v;
        }
        const int own_lo = incl - n_slots, own_hi = incl;
        sm.pref[lane] = (uint16_t)own_lo;
        const int total = __shfl_sync(FULLMASK, incl, 31);
        __syncwarp();

        // ---- phase B: one lane per (env, slot) item
        auto slow_item = [&](uint32_t packed) {
            const int env = (int)(packed >> 8), slot = (int)(packed & 0xffu);
            int ori, c;
            slot_to_placement(s_piece[sm.pid[env]], C, slot, ori, c);
            Eval ev;
            eval_slow<C, R>(sm.rec + env * K::WORDS + K::COLX + 2, s_ori[ori], c, ev, nullptr);
            if (!ev.terminal || want_terminal) emit_features(feats, e0 + env, a_stride, slot, ev, dirs);
            if (!ev.terminal) atomicOr(&acc32[2 * env + (slot >> 5)], 1u << (slot & 31));
        };
        unsigned long long vmask = 0ull;                   // owner lane: legal-slot mask of its env
        int qn = 0, cum = 0;                               // warp-uniform: queue length, envs started before base
        for (int base = 0; base < total; base += 32) {
            const uint32_t heads = window_heads(own_lo, n_slots, base);
            const int i = base + lane;
            bool slow = false, ok = false;
            uint32_t packed = 0;
            if (i < total) {
                const int env = cum + __popc(heads & lane_le) - 1;
                const int slot = i - (int)sm.pref[env];
                int ori, c;
                slot_to_placement(s_piece[sm.pid[env]], C, slot, ori, c);
                Eval ev;
                const int status = eval_fast<C, R>(sm.rec + env * K::WORDS, s_run, s_ori[ori], c, ev);
                if (status == kFastDone) {
                    emit_features(feats, e0 + env, a_stride, slot, ev, dirs);
                    ok = true;
                } else if (status == kFastClears || want_terminal) {
                    slow = true;
                    packed = (uint32_t)(env << 8 | slot);
                }
            }
            cum += __popc(heads);
            vmask |= window_bits(__ballot_sync(FULLMASK, ok), own_lo, own_hi, base);
            const unsigned bal = __ballot_sync(FULLMASK, slow);
            if (bal) {
                if (slow) sm.queue[qn + __popc(bal & (lane_le >> 1))] = (uint16_t)packed;
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) {
                    slow_item(sm.queue[lane]);
                    __syncwarp();
                    uint16_t mv = 0;
                    if (lane < qn - 32) mv = sm.queue[32 + lane];
                    __syncwarp();
                    if (lane < qn - 32) sm.queue[lane] = mv;
                    qn -= 32;
                    __syncwarp();
                }
            }
        }
        if (lane < qn) slow_item(sm.queue[lane]);
        __syncwarp();
        if (e < sv.n_env) {
            const unsigned long long v = vmask | sm.acc[lane];
            if (valid_out) valid_out[e] = v;
            if (count_out) count_out[e] = __popcll(v);
        }
        __syncwarp();
    }
}

// Afterstates with boards (compat layer / small batches): one thread per (env, slot), general path.
template <int C, int R>
__global__ void __launch_bounds__(128)
k_afterstates_export(StateView sv, float *__restrict__ feats, uint16_t *__restrict__ rows_out,
                     uint8_t *__restrict__ heights_out, int32_t *__restrict__ info_out, int a_stride)
{
    using S = Shape<C, R>;
    __shared__ uint32_t s_ori[32], s_piece[16];
    stage_tables(s_ori, s_piece);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = idx / a_stride;
    const int slot = (int)(idx % a_stride);
    if (e >= sv.n_env) return;
    const Meta mt = unpack_meta(sv.meta[e]);
    const uint32_t pw = s_piece[mt.piece];
    if (slot >= piece_num_slots(pw, C)) return;
    uint32_t col[C], nc[C];
    load_board<C, R>(sv, e, col);
    int ori, c;
    slot_to_placement(pw, C, slot, ori, c);
    Eval ev;
    eval_slow<C, R>(col, s_ori[ori], c, ev, nc);
    if (feats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) feats[idx * 8 + i] = ev.f[i];
    }
    if (rows_out) {
        uint32_t w[S::NW];
        cols_to_rows<C, R>(nc, w);
#pragma unroll
        for (int r = 0; r < S::N; ++r) rows_out[idx * S::N + r] = (uint16_t)(w[r >> 1] >> (16 * (r & 1)));
    }
    if (heights_out) {
#pragma unroll
        for (int k = 0; k < C; ++k) heights_out[idx * C + k] = (uint8_t)height_of(nc[k]);
    }
    if (info_out) {
        info_out[idx * 4 + 0] = ev.a; info_out[idx * 4 + 1] = (int32_t)ev.full;
        info_out[idx * 4 + 2] = ev.terminal; info_out[idx * 4 + 3] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// K2 step
// ---------------------------------------------------------------------------------------------
template <int C, int R>
__global__ void __launch_bounds__(128)
k_step(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, const int32_t *__restrict__ actions,
       const uint8_t *__restrict__ tape, float *__restrict__ obs, int32_t *__restrict__ reward,
       uint8_t *__restrict__ done, int32_t *__restrict__ lines, int32_t *status, int flags, F8 dirs)
{
    __shared__ uint32_t s_ori[32], s_piece[16];
    stage_tables(s_ori, s_piece);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= sv.n_env) return;
    uint32_t col[C];
    load_board<C, R>(sv, e, col);
    Meta mt = unpack_meta(sv.meta[e]);
    uint2 ep = sv.epi[e];
    const uint32_t pw = s_piece[mt.piece];
    const int n_slots = piece_num_slots(pw, C);
    const int action = actions[e];
    int sel = -1;
    if (action >= 0) {
        if (flags & TB_FLAG_ACTION_IS_SLOT) {
            if (action < n_slots) {
                int ori, c;
                slot_to_placement(pw, C, action, ori, c);
                if (placement_valid<C, R>(col, s_ori[ori], c)) sel = action;
            }
        } else {
            const unsigned long long vm = valid_mask<C, R>(col, pw, s_ori);   // game.py:69
            if (action < __popcll(vm)) sel = nth_set_bit(vm, action);         // game.py:83
        }
    }
    if (sel < 0) {                                                            // IndexError in the reference
        if (status) atomicOr(status, 1);
        return;
    }
    int ori, c;
    slot_to_placement(pw, C, sel, ori, c);
    Eval ev;
    eval_slow<C, R>(col, s_ori[ori], c, ev, col);                             // current_state = afterstates[action]
    const int lc = popc32(ev.full);                                           // game.py:85
    int rew = lc - 1;                                                         // game.py:86
    const uint64_t key = env_key(seed, (uint64_t)(env_offset + e));
    mt.piece = draw_piece(piece_set, key, mt, tape, e);                       // game.py:87
    const bool dn = !any_valid<C, R>(col, s_piece[mt.piece], s_ori);          // game.py:88,94-100
    if (dn) rew -= 100;                                                       // game.py:89-90
    ep.x += 1u; ep.y += (uint32_t)lc;
    if (obs) {
        float4 *o = reinterpret_cast<float4 *>(obs + e * 8);
        o[0] = make_float4(ev.f[0] * dirs.v[0], ev.f[1] * dirs.v[1], ev.f[2] * dirs.v[2], ev.f[3] * dirs.v[3]);
        o[1] = make_float4(ev.f[4] * dirs.v[4], ev.f[5] * dirs.v[5], ev.f[6] * dirs.v[6], ev.f[7] * dirs.v[7]);
    }
    if (reward) reward[e] = rew;
    if (done) done[e] = (uint8_t)dn;
    if (lines) lines[e] = lc;
    if (dn && (flags & TB_FLAG_AUTO_RESET) && !tape) {                        // example_play.py:20-21
#pragma unroll
        for (int k = 0; k < C; ++k) col[k] = 0u;
        mt.piece = draw_piece(piece_set, key, mt, nullptr, e);
        ep = make_uint2(0u, 0u);
    }
    store_board<C, R>(sv, e, col);
    sv.meta[e] = pack_meta<C>(col, mt);
    sv.epi[e] = ep;
}

// ---------------------------------------------------------------------------------------------
// K3 rollouts
// ---------------------------------------------------------------------------------------------
struct LaneStats {
    int placements, episodes, lines, reward, afterstates, l0, l1, l2, l3, l4, max_ep_lines, max_ep_steps;
    long long sum_ep_steps, sum_ep_lines;
};
__device__ __forceinline__ void stats_zero(LaneStats &s)
{
    s.placements = s.episodes = s.lines = s.reward = s.afterstates = 0;
    s.l0 = s.l1 = s.l2 = s.l3 = s.l4 = 0;
    s.max_ep_lines = s.max_ep_steps = 0;
    s.sum_ep_steps = s.sum_ep_lines = 0;
}
__device__ __forceinline__ long long warp_sum(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}
__device__ __forceinline__ long long warp_max(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const long long u = __shfl_xor_sync(FULLMASK, v, o); v = u > v ? u : v; }
    return v;
}
// warp -> block (shared) -> global.  Must be reached by every thread of the block.
__device__ __forceinline__ void stats_flush(const LaneStats &s, long long *s_blk, int64_t *stats)
{
    if (threadIdx.x < TB_ST_COUNT) s_blk[threadIdx.x] = 0;
    __syncthreads();
    long long v[TB_ST_COUNT];
    v[TB_ST_PLACEMENTS] = s.placements; v[TB_ST_EPISODES] = s.episodes; v[TB_ST_LINES] = s.lines;
    v[TB_ST_REWARD] = s.reward; v[TB_ST_AFTERSTATES] = s.afterstates;
    v[TB_ST_LINES0] = s.l0; v[TB_ST_LINES1] = s.l1; v[TB_ST_LINES2] = s.l2; v[TB_ST_LINES3] = s.l3; v[TB_ST_LINES4] = s.l4;
    v[TB_ST_MAX_EP_LINES] = s.max_ep_lines; v[TB_ST_MAX_EP_STEPS] = s.max_ep_steps;
    v[TB_ST_SUM_EP_STEPS] = s.sum_ep_steps; v[TB_ST_SUM_EP_LINES] = s.sum_ep_lines;
    v[TB_ST_RESERVED0] = 0; v[TB_ST_RESERVED1] = 0;
#pragma unroll
    for (int i = 0; i < TB_ST_COUNT; ++i) {
        const bool is_max = (i == TB_ST_MAX_EP_LINES || i == TB_ST_MAX_EP_STEPS);
        const long long r = is_max ? warp_max(v[i]) : warp_sum(v[i]);
        if ((threadIdx.x & 31) == 0 && r != 0) {
            if (is_max) atomicMax(&s_blk[i], r);
            else atomicAdd((unsigned long long *)&s_blk[i], (unsigned long long)r);
        }
    }
    __syncthreads();
    if (threadIdx.x < TB_ST_COUNT) {
        const int i = threadIdx.x;
        const long long r = s_blk[i];
        if (r != 0) {
            if (i == TB_ST_MAX_EP_LINES || i == TB_ST_MAX_EP_STEPS) atomicMax((long long *)&stats[i], r);
            else atomicAdd((unsigned long long *)&stats[i], (unsigned long long)r);
        }
    }
}

// Apply the chosen placement to the lane's env: lock, clear, reward, next piece, game-over, auto-reset.
template <int C, int R>
__device__ __forceinline__ void apply_placement(uint32_t *col, Meta &mt, uint2 &ep, uint32_t d, int c, int piece_set,
                                                uint64_t key, const uint32_t *s_ori, const uint32_t *s_piece,
                                                LaneStats &st)
{
    int a, term;
    uint32_t full;
    place_and_clear<C, R>(col, d, c, a, full, term);
    const int lc = popc32(full);
    int rew = lc - 1;
    mt.piece = set_piece(piece_set, bag_draw(set_size(piece_set), key, mt.bag, mt.draws));
    const bool dn = !any_valid<C, R>(col, s_piece[mt.piece], s_ori);
    if (dn) rew -= 100;
    ep.x += 1u; ep.y += (uint32_t)lc;
    st.placements += 1; st.lines += lc; st.reward += rew;
    st.l0 += (lc == 0); st.l1 += (lc == 1); st.l2 += (lc == 2); st.l3 += (lc == 3); st.l4 += (lc == 4);
    if (dn) {
        st.episodes += 1;
        st.sum_ep_steps += ep.x; st.sum_ep_lines += ep.y;
        st.max_ep_lines = imax(st.max_ep_lines, (int)ep.y);
        st.max_ep_steps = imax(st.max_ep_steps, (int)ep.x);
#pragma unroll
        for (int k = 0; k < C; ++k) col[k] = 0u;
        mt.piece = set_piece(piece_set, bag_draw(set_size(piece_set), key, mt.bag, mt.draws));
        ep = make_uint2(0u, 0u);
    }
}

// random policy: everything is per-env, one thread per env, board in registers for all n_steps
template <int C, int R>
__global__ void __launch_bounds__(128)
k_rollout_random(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, int n_steps, int64_t *stats)
{
    __shared__ uint32_t s_ori[32], s_piece[16];
    __shared__ long long s_blk[TB_ST_COUNT];
    stage_tables(s_ori, s_piece);
    LaneStats st;
    stats_zero(st);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < sv.n_env; e += (int64_t)gridDim.x * blockDim.x) {
        uint32_t col[C];
        load_board<C, R>(sv, e, col);
        Meta mt = unpack_meta(sv.meta[e]);
        uint2 ep = sv.epi[e];
        const uint64_t key = env_key(seed, (uint64_t)(env_offset + e));
        for (int t = 0; t < n_steps; ++t) {
            const uint32_t pw = s_piece[mt.piece];
            const unsigned long long vm = valid_mask<C, R>(col, pw, s_ori);
            const int nv = __popcll(vm);
            st.afterstates += piece_num_slots(pw, C);
            const int action = (int)bounded(rng32(key, mt.draws, 1u), (uint32_t)nv);
            const int slot = nth_set_bit(vm, action);
            int ori, c;
            slot_to_placement(pw, C, slot, ori, c);
            apply_placement<C, R>(col, mt, ep, s_ori[ori], c, piece_set, key, s_ori, s_piece, st);
        }
        store_board<C, R>(sv, e, col);
        sv.meta[e] = pack_meta<C>(col, mt);
        sv.epi[e] = ep;
    }
    stats_flush(st, s_blk, stats);
}

__device__ __forceinline__ uint32_t orderable(float f)
{
    const uint32_t u = __float_as_uint(f + 0.0f);          // + 0.0f: -0.0 -> +0.0 (np.argmax treats them equal)
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// greedy linear policy: warp tile of 32 envs; phases A (lane per env) / B (lane per afterstate) / C (lane per env)
template <int C, int R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_rollout_greedy(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, int n_steps, F8 wts, int64_t *stats)
{
    using K = Rec<C, R>;
    __shared__ TileSmem<C, R> s_tile[WARPS];
    __shared__ uint32_t s_ori[32], s_piece[16];
    __shared__ uint16_t s_run[RunTab<R>::SIZE];
    __shared__ long long s_blk[TB_ST_COUNT];
    stage_all<R>(s_ori, s_piece, s_run);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lane_le = (2u << lane) - 1u;
    TileSmem<C, R> &sm = s_tile[warp];
    const int64_t n_tiles = (sv.n_env + 31) >> 5;
    LaneStats st;
    stats_zero(st);

    for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < n_tiles; tile += (int64_t)gridDim.x * WARPS) {
        const int64_t e = tile * 32 + lane;
        const bool active = e < sv.n_env;
        uint32_t col[C];
        Meta mt; mt.piece = 0; mt.bag = 0u; mt.draws = 0u;
        uint2 ep = make_uint2(0u, 0u);
        uint64_t key = 0;
        if (active) {
            load_board<C, R>(sv, e, col);
            mt = unpack_meta(sv.meta[e]);
            ep = sv.epi[e];
            key = env_key(seed, (uint64_t)(env_offset + e));
        } else {
#pragma unroll
            for (int k = 0; k < C; ++k) col[k] = 0u;
        }
        for (int t = 0; t < n_steps; ++t) {
            // ---- phase A
            int n_slots = 0;
            if (active) {
                build_env<C, R>(col, sm.rec + lane * K::WORDS);
                sm.pid[lane] = (uint8_t)mt.piece;
                n_slots = piece_num_slots(s_piece[mt.piece], C);
                st.afterstates += n_slots;
            }
            sm.acc[lane] = 0ull;
            int incl = n_slots;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULLMASK, incl, o);
                if (lane >= o) incl += v;
            }
            const int own_lo = incl - n_slots, own_hi = incl;
            sm.pref[lane] = (uint16_t)own_lo;
            const int total = __shfl_sync(FULLMASK, incl, 31);
            __syncwarp();
            // ---- phase B: score every legal afterstate, keep the first arg-max per env.
            // Fast-path items: segmented max-scan over the window's lanes (items of an env are contiguous), the env's
            // owner lane picks up the result of its segment.  Slow-path items (rare): 64-bit atomicMax in shared.
            auto slow_item = [&](uint32_t packed) {
                const int env = (int)(packed >> 8), slot = (int)(packed & 0xffu);
                int ori, c;
                slot_to_placement(s_piece[sm.pid[env]], C, slot, ori, c);
                Eval ev;
                eval_slow<C, R>(sm.rec + env * K::WORDS + K::COLX + 2, s_ori[ori], c, ev, nullptr);
                if (!ev.terminal) {
                    const float score = fitness(ev.f, wts.v);                   // game.py:109-120
                    atomicMax(&sm.acc[env], ((unsigned long long)orderable(score) << 32) |
                                                (unsigned long long)(0xFFFFFFFFu - (uint32_t)slot));
                }
            };
            uint32_t best_key = 0u;                        // owner lane: best orderable score so far (0 = none)
            int best_slot = 0;
            int qn = 0, cum = 0;
            for (int base = 0; base < total; base += 32) {
                const uint32_t heads = window_heads(own_lo, n_slots, base);
                const int i = base + lane;
                bool slow = false;
                uint32_t packed = 0, key = 0u;
                int slot = 0;
                if (i < total) {
                    const int env = cum + __popc(heads & lane_le) - 1;
                    slot = i - (int)sm.pref[env];
                    int ori, c;
                    slot_to_placement(s_piece[sm.pid[env]], C, slot, ori, c);
                    Eval ev;
                    const int status = eval_fast<C, R>(sm.rec + env * K::WORDS, s_run, s_ori[ori], c, ev);
                    if (status == kFastDone) key = orderable(fitness(ev.f, wts.v));
                    else if (status == kFastClears) { slow = true; packed = (uint32_t)(env << 8 | slot); }
                }
                cum += __popc(heads);
                // lanes below this one that belong to the same env (its segment may have started in an earlier window)
                const int dist = lane - (31 - __clz((int)(heads & lane_le)));
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t k2 = __shfl_up_sync(FULLMASK, key, o);
                    const int s2 = __shfl_up_sync(FULLMASK, slot, o);
                    if (o <= dist && o <= lane && k2 >= key) { key = k2; slot = s2; }   // ties: the earlier slot wins
                }
                const int seg_s = imax(own_lo, base), seg_t = imin(own_hi, base + 32);
                const int src = seg_s < seg_t ? seg_t - 1 - base : lane;
                const uint32_t k3 = __shfl_sync(FULLMASK, key, src);
                const int s3 = __shfl_sync(FULLMASK, slot, src);
                if (seg_s < seg_t && k3 > best_key) { best_key = k3; best_slot = s3; }   // earlier windows win ties
                const unsigned bal = __ballot_sync(FULLMASK, slow);
                if (bal) {
                    if (slow) sm.queue[qn + __popc(bal & (lane_le >> 1))] = (uint16_t)packed;
                    qn += __popc(bal);
                    __syncwarp();
                    if (qn >= 32) {
                        slow_item(sm.queue[lane]);
                        __syncwarp();
                        uint16_t mv = 0;
                        if (lane < qn - 32) mv = sm.queue[32 + lane];
                        __syncwarp();
                        if (lane < qn - 32) sm.queue[lane] = mv;
                        qn -= 32;
                        __syncwarp();
                    }
                }
            }
            if (lane < qn) slow_item(sm.queue[lane]);
            __syncwarp();
            // ---- phase C
            if (active) {
                unsigned long long best = sm.acc[lane];
                if (best_key != 0u) {
                    const unsigned long long k64 = ((unsigned long long)best_key << 32) |
                                                   (unsigned long long)(0xFFFFFFFFu - (uint32_t)best_slot);
                    best = k64 > best ? k64 : best;
                }
                if (best != 0ull) {
                    const int slot = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
                    int ori, c;
                    slot_to_placement(s_piece[mt.piece], C, slot, ori, c);
                    apply_placement<C, R>(col, mt, ep, s_ori[ori], c, piece_set, key, s_ori, s_piece, st);
                } else {
                    // no legal placement: only reachable from a caller-supplied dead state -> start a new episode
#pragma unroll
                    for (int k = 0; k < C; ++k) col[k] = 0u;
                    mt.piece = set_piece(piece_set, bag_draw(set_size(piece_set), key, mt.bag, mt.draws));
                    ep = make_uint2(0u, 0u);
                }
            }
            __syncwarp();
        }
        if (active) {
            store_board<C, R>(sv, e, col);
            sv.meta[e] = pack_meta<C>(col, mt);
            sv.epi[e] = ep;
        }
    }
    stats_flush(st, s_blk, stats);
}

// ---------------------------------------------------------------------------------------------
// state interchange + State evaluation on caller boards
// ---------------------------------------------------------------------------------------------
template <int C, int R>
__global__ void k_export(StateView sv, int64_t first, int64_t count, uint16_t *__restrict__ rows_out,
                         uint8_t *__restrict__ heights_out, uint8_t *__restrict__ piece_out)
{
    using S = Shape<C, R>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t e = first + i;
    if (rows_out) {
#pragma unroll
        for (int b = 0; b < S::NB; ++b) {
            const uint4 v = sv.planes[(int64_t)b * sv.n_env + e];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = 8 * b + k;
                if (r < S::N) rows_out[i * S::N + r] = (uint16_t)(w[k >> 1] >> (16 * (k & 1)));
            }
        }
    }
    const uint4 m = sv.meta[e];
    if (heights_out) {
        const uint32_t w[3] = {m.x, m.y, m.z};
#pragma unroll
        for (int c = 0; c < C; ++c) heights_out[i * C + c] = (uint8_t)((w[c >> 2] >> (8 * (c & 3))) & 0xffu);
    }
    if (piece_out) piece_out[i] = (uint8_t)((m.z >> 16) & 0xffu);
}

template <int C, int R>
__global__ void k_import(StateView sv, int64_t first, int64_t count, const uint16_t *__restrict__ rows_in,
                         const uint8_t *__restrict__ piece_in)
{
    using S = Shape<C, R>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t e = first + i;
    uint32_t w[S::NW];
#pragma unroll
    for (int k = 0; k < S::NW; ++k) w[k] = 0u;
#pragma unroll
    for (int r = 0; r < S::N; ++r) w[r >> 1] |= ((uint32_t)rows_in[i * S::N + r] & S::FULLROW) << (16 * (r & 1));
    uint32_t col[C];
    rows_to_cols<C, R>(w, col);
#pragma unroll
    for (int b = 0; b < S::NB; ++b)
        sv.planes[(int64_t)b * sv.n_env + e] = make_uint4(w[4 * b], w[4 * b + 1], w[4 * b + 2], w[4 * b + 3]);
    Meta mt = unpack_meta(sv.meta[e]);
    if (piece_in) mt.piece = piece_in[i];
    sv.meta[e] = pack_meta<C>(col, mt);
}

template <int C, int R>
__global__ void k_eval_states(int64_t n, const uint16_t *__restrict__ rows_in, const int32_t *__restrict__ params,
                              uint16_t *__restrict__ rows_out, uint8_t *__restrict__ heights_out,
                              int32_t *__restrict__ info_out, float *__restrict__ feats_out)
{
    using S = Shape<C, R>;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[S::NW];
#pragma unroll
    for (int k = 0; k < S::NW; ++k) w[k] = 0u;
#pragma unroll
    for (int r = 0; r < S::N; ++r) w[r >> 1] |= ((uint32_t)rows_in[i * S::N + r] & S::FULLROW) << (16 * (r & 1));
    uint32_t col[C];
    rows_to_cols<C, R>(w, col);
    int a = 0, chg = 1, bonus2 = 0;
    uint32_t ppcr = 0;
    if (params) {
        a = params[i * 4 + 0]; chg = params[i * 4 + 1]; ppcr = (uint32_t)params[i * 4 + 2]; bonus2 = params[i * 4 + 3];
        a = imin(imax(a, 0), S::N - 1); chg = imin(imax(chg, 0), 4);
    }
    Eval ev;
    eval_state<C, R>(col, a, chg, ppcr, bonus2, ev);
    if (rows_out) {
        cols_to_rows<C, R>(col, w);
#pragma unroll
        for (int r = 0; r < S::N; ++r) rows_out[i * S::N + r] = (uint16_t)(w[r >> 1] >> (16 * (r & 1)));
    }
    if (heights_out) {
#pragma unroll
        for (int k = 0; k < C; ++k) heights_out[i * C + k] = (uint8_t)height_of(col[k]);
    }
    if (info_out) {
        info_out[i * 4 + 0] = popc32(ev.full); info_out[i * 4 + 1] = (int32_t)ev.full;
        info_out[i * 4 + 2] = ev.terminal; info_out[i * 4 + 3] = 0;
    }
    if (feats_out) {
#pragma unroll
        for (int k = 0; k < 8; ++k) feats_out[i * 8 + k] = ev.f[k];
    }
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

}  // namespace tb

// =============================================================================================
// C ABI
// =============================================================================================
using namespace tb;

// board shapes compiled in: the three of BASELINE.json's configs plus two extras (mid-size, tiny edge case)
#define TB_SHAPES(X) X(10, 20) X(10, 10) X(6, 12) X(8, 16) X(4, 4)

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
static unsigned grid_for(int64_t work_items, int per_block, int blocks_per_sm)
{
    const int64_t need = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * blocks_per_sm;         // persistent: a multiple of the SM count
    return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}
static F8 f8_from(const float *p, float dflt)
{
    F8 r;
    for (int i = 0; i < 8; ++i) r.v[i] = p ? p[i] : dflt;
    return r;
}

extern "C" {

int tb_version(void) { return TB_VERSION; }
const char *tb_last_error(void) { return g_err; }

int tb_supported_shape(int C, int R)
{
#define X(c, r) if (C == c && R == r) return 1;
    TB_SHAPES(X)
#undef X
    return 0;
}

size_t tb_state_bytes(int C, int R, int64_t n_env)
{
#define X(c, r) if (C == c && R == r) return (size_t)n_env * (size_t)(16 * (Shape<c, r>::NB + 1) + 8);
    TB_SHAPES(X)
#undef X
    return 0;
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
    if (!tb_supported_shape(C, R)) return fail("%s: unsupported board shape", __func__);

int tb_reset(void *state, int C, int R, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
             const uint8_t *piece_tape, const uint8_t *reset_mask, void *stream)
{
    TB_CHECK_COMMON();
    if (piece_set < 0 || piece_set > 1) return fail("%s: piece_set must be 0 or 1", __func__);
    cudaStream_t st = (cudaStream_t)stream;
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        k_reset<c, r><<<(unsigned)((n_env + 255) / 256), 256, 0, st>>>(make_view<c, r>(state, n_env),    \
            env_offset, seed, piece_set, piece_tape, reset_mask);                                        \
        return check_launch("tb_reset");                                                                 \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_afterstates(const void *state, int C, int R, int64_t n_env, float *feats_out, uint64_t *valid_out,
                   int32_t *count_out, int a_stride, const float *directions, int flags, void *stream)
{
    TB_CHECK_COMMON();
    if (!feats_out) return fail("%s: feats_out is required", __func__);
    if (a_stride < 1) return fail("%s: a_stride must be >= the piece set's slot count", __func__);
    cudaStream_t st = (cudaStream_t)stream;
    const F8 dirs = f8_from(directions, 1.0f);
    constexpr int WARPS = 4;
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        k_afterstates<c, r, WARPS><<<grid_for((n_env + 31) / 32, WARPS, 8), WARPS * 32, 0, st>>>(        \
            make_view<c, r>(state, n_env), feats_out, (unsigned long long *)valid_out, count_out, a_stride, dirs, \
            flags);                                                                                      \
        return check_launch("tb_afterstates");                                                           \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_afterstates_export(const void *state, int C, int R, int64_t n_env, float *feats_out, uint16_t *rows_out,
                          uint8_t *heights_out, int32_t *info_out, int a_stride, void *stream)
{
    TB_CHECK_COMMON();
    if (a_stride < 1) return fail("%s: a_stride must be >= the piece set's slot count", __func__);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = n_env * a_stride;
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        k_afterstates_export<c, r><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(                     \
            make_view<c, r>(state, n_env), feats_out, rows_out, heights_out, info_out, a_stride);        \
        return check_launch("tb_afterstates_export");                                                    \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_step(void *state, int C, int R, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
            const int32_t *actions, const uint8_t *piece_tape, float *obs_out, int32_t *reward_out, uint8_t *done_out,
            int32_t *lines_out, int32_t *status_out, int flags, void *stream)
{
    TB_CHECK_COMMON();
    if (!actions) return fail("%s: actions is required", __func__);
    if (piece_set < 0 || piece_set > 1) return fail("%s: piece_set must be 0 or 1", __func__);
    cudaStream_t st = (cudaStream_t)stream;
    const F8 dirs = f8_from(nullptr, 1.0f);
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        k_step<c, r><<<(unsigned)((n_env + 127) / 128), 128, 0, st>>>(make_view<c, r>(state, n_env),     \
            env_offset, seed, piece_set, actions, piece_tape, obs_out, reward_out, done_out, lines_out,  \
            status_out, flags, dirs);                                                                    \
        return check_launch("tb_step");                                                                  \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_rollout(void *state, int C, int R, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
               int n_steps, int policy, const float *weights, int64_t *stats, void *stream)
{
    TB_CHECK_COMMON();
    if (!stats) return fail("%s: stats is required", __func__);
    if (n_steps < 0) return fail("%s: n_steps must be >= 0", __func__);
    if (piece_set < 0 || piece_set > 1) return fail("%s: piece_set must be 0 or 1", __func__);
    if (policy == TB_POLICY_GREEDY && !weights) return fail("%s: greedy policy needs weights", __func__);
    if (policy != TB_POLICY_GREEDY && policy != TB_POLICY_RANDOM) return fail("%s: unknown policy", __func__);
    cudaStream_t st = (cudaStream_t)stream;
    const F8 wts = f8_from(weights, 0.0f);
    constexpr int WARPS = 4;
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        if (policy == TB_POLICY_RANDOM)                                                                  \
            k_rollout_random<c, r><<<grid_for(n_env, 128, 8), 128, 0, st>>>(make_view<c, r>(state, n_env), \
                env_offset, seed, piece_set, n_steps, stats);                                            \
        else                                                                                             \
            k_rollout_greedy<c, r, WARPS><<<grid_for((n_env + 31) / 32, WARPS, 4), WARPS * 32, 0, st>>>( \
                make_view<c, r>(state, n_env), env_offset, seed, piece_set, n_steps, wts, stats);        \
        return check_launch("tb_rollout");                                                               \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_export_boards(const void *state, int C, int R, int64_t n_env, int64_t first, int64_t count,
                     uint16_t *rows_out, uint8_t *heights_out, uint8_t *piece_out, void *stream)
{
    TB_CHECK_COMMON();
    if (first < 0 || count < 0 || first + count > n_env) return fail("%s: env range out of bounds", __func__);
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        k_export<c, r><<<(unsigned)((count + 127) / 128), 128, 0, st>>>(make_view<c, r>(state, n_env),   \
            first, count, rows_out, heights_out, piece_out);                                             \
        return check_launch("tb_export_boards");                                                         \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_import_boards(void *state, int C, int R, int64_t n_env, int64_t first, int64_t count,
                     const uint16_t *rows_in, const uint8_t *piece_in, void *stream)
{
    TB_CHECK_COMMON();
    if (first < 0 || count < 0 || first + count > n_env) return fail("%s: env range out of bounds", __func__);
    if (!rows_in) return fail("%s: rows_in is required", __func__);
    if (count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        k_import<c, r><<<(unsigned)((count + 127) / 128), 128, 0, st>>>(make_view<c, r>(state, n_env),   \
            first, count, rows_in, piece_in);                                                            \
        return check_launch("tb_import_boards");                                                         \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_eval_states(int C, int R, int64_t n, const uint16_t *rows_in, const int32_t *params, uint16_t *rows_out,
                   uint8_t *heights_out, int32_t *info_out, float *feats_out, void *stream)
{
    if (n <= 0) return fail("%s: n must be positive", __func__);
    if (!tb_supported_shape(C, R)) return fail("%s: unsupported board shape", __func__);
    if (!rows_in) return fail("%s: rows_in is required", __func__);
    cudaStream_t st = (cudaStream_t)stream;
#define X(c, r)                                                                                          \
    if (C == c && R == r) {                                                                              \
        k_eval_states<c, r><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, rows_in, params, rows_out,   \
            heights_out, info_out, feats_out);                                                           \
        return check_launch("tb_eval_states");                                                           \
    }
    TB_SHAPES(X)
#undef X
    return -1;
}

int tb_fitness(int64_t n, const float *feats, const float *weights, float *out, void *stream)
{
    if (n <= 0) return fail("%s: n must be positive", __func__);
    if (!feats || !weights || !out) return fail("%s: null argument", __func__);
    if ((reinterpret_cast<uintptr_t>(feats) & 15u) != 0) return fail("%s: feats must be 16-byte aligned", __func__);
    k_fitness<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, feats, f8_from(weights, 0.0f), out);
    return check_launch("tb_fitness");
}

}  // extern "C"
