// tb_kernels.cuh -- sm_100a kernel templates of the batched Tetris environment, per board shape <C, R>.
//
// Kernels:
//   k_reset        K0  Tetris.__init__/reset            game.py:21-63
//   k_afterstates  K1  Tetris.get_after_states           game.py:67-80 -> tetromino.py -> state.py
//   k_step         K2  Tetris.step + is_game_over        game.py:82-100
//   k_rollout_*    K3  example_play.py:11-21 loop fused with an in-kernel policy
//
// Mapping.  The work is integer/bit manipulation on a few dozen bytes per env -- no GEMM shape anywhere, so
// no tensor cores.  Envs are independent.  K1 / K3 give a CTA a tile of 256 envs: per-env work (load + transpose,
// the env record, applying a placement, RNG, game-over) runs one thread per env; per-afterstate work runs in
// windows of envs that hold the same piece, lane = (env of the window, anchor column), so the piece's orientation
// loop and width are warp-uniform and lanes stay busy although pieces have 9..34 placements.  Placements that
// clear a line are rare and costly (from-scratch evaluation): they are marked in per-env bit masks and evaluated
// 32 at a time instead of diverging the common incremental path.  See DESIGN.md section 3.
//
// Build layout: this header is compiled once per board shape by tb_shape.cu (-DTB_C=.. -DTB_R=..), which exports the
// shape's launchers as a table of function pointers (ShapeVT, tb_shape.h); tb_abi.cu holds the extern "C" entry points
// of include/tetris_b200.h and dispatches on the shape.  One translation unit per shape keeps a rebuild at the cost of
// its slowest shape and lets further shapes be compiled on demand (tb_load_shape).
//
// HBM layout: see include/tetris_b200.h (row masks, 8 rows per 128-bit word, SoA over envs -> every global
// load/store of state is a coalesced 128-bit access).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/tetris_b200.h"
#include "tb_core.cuh"
#include "tb_shape.h"

namespace tb {

static __constant__ uint32_t c_ori[kNumOris] = { TB_ORI_TABLE(TB_X_ORI) };
static __constant__ uint32_t c_piece[kNumPieces] = { TB_PIECE_TABLE(TB_X_PIECE) };

constexpr unsigned FULLMASK = 0xFFFFFFFFu;

struct StateView {
    uint4 *planes;      // [NB][n_env]
    uint4 *meta;        // [n_env]
    uint2 *epi;         // [n_env]
    int64_t n_env;
};

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
// meta word: bytes 0..9 column heights (boards of up to 10 columns; wider boards leave them zero -- heights are always
// derivable from the row planes, which is what tb_export_boards does), byte 10 piece, byte 11 bag, bytes 12..15 draws
template <int C>
__device__ __forceinline__ uint4 pack_meta(const uint32_t *col, Meta mt)
{
    uint32_t w[3] = {0u, 0u, 0u};
    if (C <= 10) {
#pragma unroll
        for (int c = 0; c < C; ++c) w[(c >> 2) % 3] |= (uint32_t)height_of(col[c]) << (8 * (c & 3));
    }
    return make_uint4(w[0], w[1], w[2] | ((uint32_t)mt.piece << 16) | (mt.bag << 24), mt.draws);
}
template <int C>
__device__ __forceinline__ int max_height_raw(const uint32_t *col)
{
    uint32_t any = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) any |= col[c];
    return height_of(any);
}
template <int C, int R>
__device__ __forceinline__ void load_board(const StateView &sv, int64_t e, uint32_t *col)
{
    using S = Shape<C, R>;
    TB_CHECK(e >= 0 && e < sv.n_env);
    uint32_t w[S::NW];
#pragma unroll
    for (int b = 0; b < S::NB; ++b) {
        const uint4 v = sv.planes[(int64_t)b * sv.n_env + e];
        w[4 * b] = v.x; w[4 * b + 1] = v.y; w[4 * b + 2] = v.z; w[4 * b + 3] = v.w;
    }
    rows_to_cols<C, R>(w, col);
    TB_CHECK(max_height_raw<C>(col) <= S::N);
}
template <int C, int R>
__device__ __forceinline__ void store_board(const StateView &sv, int64_t e, const uint32_t *col)
{
    using S = Shape<C, R>;
    TB_CHECK(e >= 0 && e < sv.n_env);
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
// A tape id that names no piece makes the env inert (kPieceVoid: no afterstates, never stepped) instead of indexing
// the piece tables out of bounds.
__device__ __forceinline__ int checked_piece(int id)
{
    return (id < kNumPieces || id == kPieceDead) ? id : kPieceVoid;
}
__device__ __forceinline__ int draw_piece(int piece_set, uint64_t key, Meta &mt, const uint8_t *tape, int64_t e)
{
    if (tape) { mt.draws += 1; return checked_piece((int)tape[e]); }
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
// mask of legal slots (game.py:69); 0 <=> the piece cannot be placed: game over (is_game_over, game.py:94-100)
template <int C, int R>
__device__ __forceinline__ unsigned long long valid_mask(const uint32_t *col, uint32_t pw, const uint32_t *s_ori)
{
    return valid_slots<C, R>(col, pw, s_ori);
}
template <int C, int R>
__device__ __forceinline__ bool any_valid(const uint32_t *col, uint32_t pw, const uint32_t *s_ori)
{
    return valid_slots<C, R, true>(col, pw, s_ori) != 0ull;
}
// position of the n-th (0-based) set bit of m; n < popc(m)
__device__ __forceinline__ int nth_set_bit(unsigned long long m, int n)
{
    TB_CHECK(n >= 0 && n < __popcll(m));
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    const int plo = __popc(lo);
    const bool low = n < plo;
    return (low ? 0 : 32) + (int)__fns(low ? lo : hi, 0u, (low ? n : n - plo) + 1);
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
// K1 afterstates.
//
// A CTA owns a tile of TILE envs (template parameter; THREADS = TILE threads in the throughput configuration, a multiple
// of TILE in the small-batch ones).
//   Phase A (thread per env): counting sort of the tile's envs by piece -- the env's rank within its piece by a
//     shared atomic, barrier, slot = (envs with smaller pieces) + rank -- then load + transpose the board and build the
//     env record in shared memory AT ITS SLOT.
//   Phase B (warp per window): all lanes of a warp work on envs holding the SAME piece -- lane = (env k of the window,
//     anchor column c) -- and walk the piece's orientations in a warp-uniform loop, so the orientation descriptor is
//     uniform, every loop over the piece's columns has a uniform trip count, and the neighbourhood loads are shared by the
//     orientations of a column loop.  Because same-piece envs sit in consecutive slots, a window's records are at fixed
//     distances, chosen (window_stride) so that they start on disjoint ranges of shared-memory banks: the window's loads
//     are conflict-free (round 1 indexed records by env id: 37 % of K1's shared wavefronts were conflict replays).
//   Phase S: placements that clear a line only need the general evaluation for their FEATURES (legality follows from the
//     full-row count); they are marked in per-slot bit masks in phase B and evaluated one lane per item.
// ---------------------------------------------------------------------------------------------
constexpr int kNumJobs = 2 * kNumPieces;
constexpr int kNoSlot = 0xFFFF;

template <int C, int R, int TILE>
struct CtaSmem {
    using K = Env<C, R>;
    static_assert(TILE <= 256, "slot -> env map is bytes");
    uint32_t rec[TILE * K::WORDS];       // env records by SLOT (the tile's envs sorted by piece)
    alignas(16) uint32_t odesc[kNumOris][kOriWords]; // orientation descriptors, decoded (OriU): broadcast 128-bit loads
    uint32_t vloc[TILE][2];              // by slot.  K1: legal placements per column loop (16 bits per orientation)
                                         //           K3: best orderable score per column loop
    uint32_t sloc[TILE][2];              // by slot: placements that need the general evaluation (they clear a line),
                                         // per column loop, same bit layout as vloc in K1 (16 bits per orientation)
    uint16_t slot_of[TILE];              // env (thread of the tile) -> slot; kNoSlot: the env takes no part
    uint16_t sprefix[TILE];              // K1 phase S: inclusive count of slow items over the slots of each 32-slot group
    int wtot[TILE / 32];                 //             and the groups' totals
    uint8_t env_of[TILE];                // slot -> env (thread of the tile)
    uint8_t pid[TILE];                   // by slot: piece
    int cnt[2][kNumPieces + 1];          // envs per piece, double-buffered by tile / step parity
    uint32_t job[kNumJobs];              // (piece, column loop): p | l << 4 | w << 5 | n << 8 | ob << 10 | sbase << 16
    uint32_t ori[32], piece[16];
};
static_assert(sizeof(CtaSmem<10, 20, 256>) <= 57 * 1024, "K1 / K3 at 10x20: four 256-env CTAs per SM fit in shared memory");

// K3 only: best (score, slot) key among the line-clearing placements, and per-warp episode statistics
template <int TILE, int THREADS = TILE> struct BestSmem {
    unsigned long long best[TILE];       // by slot
    long long wstat[THREADS / 32][TB_ST_COUNT];
    uint8_t bslot[TILE][2];              // by slot: enumeration slot of the best score per column loop
};

// Compile-time images of the two shared-memory tables (tb_core.cuh: make_odesc_image / make_run_image, checked against
// decode_ori / run_tab_entry on the CPU by tests/hostcheck), copied -- not computed -- by every CTA: with multi-wave
// grids a CTA often handles a single tile, and building the run table in the kernel cost about 3 % of a K1 tile.
static __device__ const OdescImage g_odesc = make_odesc_image();
// The same image and the job table in CONSTANT memory, for K1's phase B: its descriptor fields and job words then come from
// the constant cache (LDC) instead of shared memory, whose pipe K1 needs for the env records and the run table
// (K1 -2.8 %; K3, measured the same way, +3.3 %: it keeps the shared-memory copies.  profiles/README.md, r2t).
static __constant__ OdescImage c_odesc = make_odesc_image();
struct JobImage { uint32_t v[kNumJobs]; };
template <int C>
constexpr JobImage make_job_image()
{
    JobImage t{};
    for (int i = 0; i < kNumJobs; ++i) {
        const int p = i >> 1, l = i & 1;
        const uint32_t pw = kPieceHost[p];
        const int n0 = pw & 3, w0 = (pw >> 2) & 7, n1 = (pw >> 5) & 3, w1 = (pw >> 7) & 7, obase = (pw >> 10) & 63;
        const int n = l ? n1 : n0, w = l ? w1 : w0, ob = l ? obase + n0 : obase, sbase = l ? n0 * (C - w0 + 1) : 0;
        t.v[i] = (uint32_t)p | (uint32_t)l << 4 | (uint32_t)w << 5 | (uint32_t)(w <= C ? n : 0) << 8 |
                 (uint32_t)ob << 10 | (uint32_t)sbase << 16;
    }
    return t;
}
template <int C> static __constant__ JobImage c_job = make_job_image<C>();
template <int R> static __device__ const RunImage<R> g_run = make_run_image<R>();

// The run-sum table must sit at a shared-memory ADDRESS that is a multiple of its size (run_sum_acc forms entry addresses
// with OR).  The shared window of a CTA does not start at a multiple of 4 KB (the first 1 KB is the system's), so the
// alignment of a static array says nothing about its address: the table gets twice its size and is placed at run time.
template <int R>
__device__ __forceinline__ uint32_t *run_tab_place(uint32_t *raw)
{
    constexpr uint32_t BYTES = (uint32_t)RunTab<R>::SIZE * 4u;
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(raw);
    return raw + (((0u - a) & (BYTES - 1u)) >> 2);
}

// copy the table image (128-bit words where the size allows); the caller synchronises
template <int R, bool PERM = true>
__device__ __forceinline__ void stage_run_table(uint32_t *s_run)
{
    constexpr int N = PERM ? RunTab<R>::WORDS : RunTab<R>::SIZE;   // with or without the permuted second copy
    if (N % 4 == 0) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_run<R>.v);
        uint4 *dst = reinterpret_cast<uint4 *>(s_run);
        for (int i = threadIdx.x; i < N / 4; i += blockDim.x) dst[i] = src[i];
    } else {
        for (int m = threadIdx.x; m < N; m += blockDim.x) s_run[m] = g_run<R>.v[m];
    }
}

// s_run: the run-sum table, placed by run_tab_place inside a static shared-memory array (run_sum_acc then forms an entry's
// address with one LOP3).  odesc / job: K3 reads them from shared memory, K1 from constant memory (c_odesc / c_job).
template <int C, int R, int TILE>
__device__ __forceinline__ void stage_cta(CtaSmem<C, R, TILE> &sm, uint32_t *s_run)
{
    if (threadIdx.x < kNumOris) sm.ori[threadIdx.x] = c_ori[threadIdx.x];
    {
        static_assert(sizeof(OdescImage) % 16 == 0, "odesc is copied in 128-bit words");
        const uint4 *src = reinterpret_cast<const uint4 *>(g_odesc.w);
        uint4 *dst = reinterpret_cast<uint4 *>(sm.odesc);
        for (int i = threadIdx.x; i < (int)(sizeof(OdescImage) / 16); i += blockDim.x) dst[i] = src[i];
    }
    if (threadIdx.x < kNumPieces) sm.piece[threadIdx.x] = c_piece[threadIdx.x];
    if (threadIdx.x < kNumJobs) {
        const int p = threadIdx.x >> 1, l = threadIdx.x & 1;
        const uint32_t pw = c_piece[p];
        const int n0 = pw & 3, w0 = (pw >> 2) & 7, n1 = (pw >> 5) & 3, w1 = (pw >> 7) & 7, obase = (pw >> 10) & 63;
        const int n = l ? n1 : n0, w = l ? w1 : w0, ob = l ? obase + n0 : obase, sbase = l ? n0 * (C - w0 + 1) : 0;
        const uint32_t jb = (uint32_t)p | (uint32_t)l << 4 | (uint32_t)w << 5 | (uint32_t)(w <= C ? n : 0) << 8 |
                            (uint32_t)ob << 10 | (uint32_t)sbase << 16;
        sm.job[threadIdx.x] = jb;
    }
    if (threadIdx.x <= kNumPieces) { sm.cnt[0][threadIdx.x] = 0; sm.cnt[1][threadIdx.x] = 0; }
    stage_run_table<R>(s_run);
    __syncthreads();
}

// Position (within a piece's run of slots) of env k of window `win`.  The records of a window's EPW envs must start on
// disjoint ranges of shared-memory banks, NC banks each (lane c of an env reads word c + const of its record): with a
// record stride of `words` (odd), envs at positions D apart are D * words banks apart.  window_stride() finds the
// smallest such D; windows are then taken from groups of EPW * D consecutive positions (window r of a group = positions
// r, r + D, .., r + (EPW-1) D).  The ragged end of a piece's run falls back to consecutive positions (a few conflicts,
// no idle windows).  10x20: records of 43 words; 3 envs x 9..10 columns -> D = 1 (11 banks apart), 4 envs x 7..8
// columns -> D = 8 (8 banks apart).
constexpr int window_stride(int epw, int nc, int words)
{
    for (int d = 1; d <= 16; ++d) {
        bool ok = true;
        for (int i = 0; i < epw && ok; ++i)
            for (int j = i + 1; j < epw && ok; ++j) {
                const int diff = ((j - i) * d * words) % 32;
                if ((diff < 32 - diff ? diff : 32 - diff) < nc) ok = false;
            }
        if (ok) return d;
    }
    return 1;
}
template <int EPW, int NC, int WORDS> struct WinStride { static constexpr int value = window_stride(EPW, NC, WORDS); };
template <int EPW, int D>
__device__ __forceinline__ int window_position(int win, int k, int np)
{
    if (D == 1) return win * EPW + k;
    constexpr int G = EPW * D;
    const int g = win / D, r = win - g * D, b = g * G;
    return (b + G <= np) ? b + r + D * k : b + EPW * r + k;
}

// Phase S work distribution.  Every lane owns one env of the warp's 32 and holds that env's slow-placement masks
// (m0, m1: column loops 0 / 1, bit 16 * o + c).  The set bits of all 32 envs are flattened over the lanes, 32 items per
// pass, so the general evaluation runs with full warps however the items are spread over the envs.
// f(owner lane, the owner's `payload`, loop l, orientation o in the loop, column c) is called for every item by one lane.
// With several warps per 32-env group (small-batch configurations) warp `sub` of `nsub` takes every nsub-th pass.
template <typename F>
__device__ __forceinline__ void for_each_slow_item(uint32_t m0, uint32_t m1, int payload, int lane, int sub, int nsub, F &&f)
{
    const int cnt = __popc(m0) + __popc(m1);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULLMASK, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(FULLMASK, incl, 31);
    for (int base = 32 * sub; base < total; base += 32 * nsub) {
        const int i = base + lane;
        int owner = 0;                                     // first lane whose inclusive prefix exceeds i
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const int v = __shfl_sync(FULLMASK, incl, owner + step - 1);
            if (v <= i) owner += step;
        }
        owner &= 31;
        const int before = __shfl_sync(FULLMASK, incl - cnt, owner);
        const uint32_t om0 = __shfl_sync(FULLMASK, m0, owner), om1 = __shfl_sync(FULLMASK, m1, owner);
        const int opay = __shfl_sync(FULLMASK, payload, owner);
        if (i < total) {
            const int r = i - before, p0 = __popc(om0);
            const int l = r >= p0;
            const int b = (int)__fns(l ? om1 : om0, 0u, (l ? r - p0 : r) + 1);
            f(owner, opay, l, b >> 4, b & 15);
        }
    }
}

// vloc / sloc words hold, per orientation, the warp ballot shifted down to the env's first lane and cut to 16 bits by the
// store: bits at and above the loop's column count belong to the next env of the window.  Readers mask with this.
template <int C>
__device__ __forceinline__ uint32_t loc_mask(uint32_t pw, int l)
{
    const int w = (int)((pw >> (l ? 7 : 2)) & 7u);
    return ((1u << (C - w + 1)) - 1u) * 0x00010001u;
}
// legal columns per orientation (16 bits each, orientation-major) -> loop-local slot bits c * n + o
__device__ __forceinline__ uint32_t slots_of(uint32_t v, int n)
{
    return n == 2 ? (spread16(v & 0xFFFFu) | (spread16(v >> 16) << 1)) : (v & 0xFFFFu);
}

// Output formats of K1 (template parameter FMT):
//   0  float32[8] per afterstate (32 B)                       1  float32[8] x feature_directions (state.py:49-50)
//   2  int16[8] = 2 x feature x direction (16 B, TB_FLAG_FEATS_I16): every feature is a half-integer below 2^14, so the
//      doubled value is an exact small integer -- half the bytes over HBM and PCIe for a host-side policy
constexpr int kFmtF32 = 0, kFmtF32Dirs = 1, kFmtI16 = 2;
// int16 pair (2 f0 d0, 2 f1 d1) without the conversion unit: 2 f d + 1.5 * 2^23 is exact (a small integer), and the
// float's low 16 mantissa bits then ARE the two's-complement int16; one FFMA per value and one PRMT per pair.
__device__ __forceinline__ uint32_t pack_i16x2(float f0, float d0, float f1, float d1)
{
    const uint32_t a = __float_as_uint(__fmaf_rn(f0, d0 + d0, 12582912.0f));
    const uint32_t b = __float_as_uint(__fmaf_rn(f1, d1 + d1, 12582912.0f));
    return __byte_perm(a, b, 0x5410);
}
template <int FMT>
__device__ __forceinline__ void emit_row(float *__restrict__ row, const Eval &ev, const F8 &dirs)
{
    if (FMT == kFmtI16) {
        *reinterpret_cast<uint4 *>(row) = make_uint4(
            pack_i16x2(ev.f[0], dirs.v[0], ev.f[1], dirs.v[1]), pack_i16x2(ev.f[2], dirs.v[2], ev.f[3], dirs.v[3]),
            pack_i16x2(ev.f[4], dirs.v[4], ev.f[5], dirs.v[5]), pack_i16x2(ev.f[6], dirs.v[6], ev.f[7], dirs.v[7]));
        return;
    }
    float4 *dst = reinterpret_cast<float4 *>(row);
    if (FMT == kFmtF32Dirs) {
        dst[0] = make_float4(ev.f[0] * dirs.v[0], ev.f[1] * dirs.v[1], ev.f[2] * dirs.v[2], ev.f[3] * dirs.v[3]);
        dst[1] = make_float4(ev.f[4] * dirs.v[4], ev.f[5] * dirs.v[5], ev.f[6] * dirs.v[6], ev.f[7] * dirs.v[7]);
    } else {
        dst[0] = make_float4(ev.f[0], ev.f[1], ev.f[2], ev.f[3]);
        dst[1] = make_float4(ev.f[4], ev.f[5], ev.f[6], ev.f[7]);
    }
}
// row of afterstate `slot` of env: 8 floats, or 8 int16 (= 4 floats' worth of bytes) in the compact format
template <int FMT>
__device__ __forceinline__ float *feat_row(float *__restrict__ feats, int64_t env, int a_stride, int slot)
{
    return feats + ((size_t)env * (size_t)a_stride + (size_t)slot) * (FMT == kFmtI16 ? 4 : 8);
}

template <int V> struct IntC { static constexpr int value = V; };

// slot of an env with piece `piece` and rank `rank` among the tile's envs of that piece: envs with smaller pieces first
__device__ __forceinline__ int slot_from_rank(const int *cnt, int piece, int rank)
{
    int base = 0;
#pragma unroll
    for (int q = 0; q < kNumPieces - 1; ++q) base += q < piece ? cnt[q] : 0;
    return base + rank;
}

// THREADS = CTA size, a multiple of TILE.  THREADS == TILE is the throughput configuration (thread per env in the
// per-env phases).  THREADS > TILE is the small-batch configuration: the per-env phases use the first TILE threads,
// phase B spreads the windows over all warps and WPG = THREADS / TILE warps share the phase-S items of each 32-env
// group, which shortens the critical path of a tile when there are fewer tiles than SMs.
template <int C, int R, int FMT, int TILE, int THREADS>
__device__ __forceinline__ void
afterstates_body(const StateView &sv, float *__restrict__ feats, unsigned long long *__restrict__ valid_out,
                 int *__restrict__ count_out, int a_stride, const F8 &dirs, int flags)
{
    using K = Env<C, R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(TILE % 32 == 0 && THREADS % TILE == 0, "TILE: whole warps of envs; THREADS: a multiple of TILE");
    constexpr int NWARPS = THREADS / 32, NGROUPS = TILE / 32;
    CtaSmem<C, R, TILE> &sm = *reinterpret_cast<CtaSmem<C, R, TILE> *>(smem_raw);
    __shared__ __align__(16) uint32_t s_run_raw[RunTab<R>::SIZE + RunTab<R>::WORDS];
    uint32_t *const s_run = run_tab_place<R>(s_run_raw);
    stage_cta(sm, s_run);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool env_thread = THREADS == TILE || tid < TILE;   // this thread owns env `tid` of the tile in the per-env phases
    const bool want_terminal = (flags & TB_FLAG_INCLUDE_TERMINAL) != 0;
    const int64_t n_tiles = (sv.n_env + TILE - 1) / TILE;
    int par = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, par ^= 1) {
        const int64_t e0 = tile * TILE, e = e0 + tid;
        int *cnt = sm.cnt[par];
        // ---- phase A1 (thread per env): rank of the env among the tile's envs with the same piece.
        // cnt[par] was zeroed during the previous tile's phase A2.
        int piece = kPieceVoid, rank = 0;
        if (env_thread && e < sv.n_env) {
            piece = unpack_meta(sv.meta[e]).piece;
            if (piece < kNumPieces) rank = atomicAdd(&cnt[piece], 1);   // finished forks (0xFF / 0xFE): no afterstates
        }
        __syncthreads();                                   // counts complete; every warp is done with the previous tile
        if (tid <= kNumPieces) sm.cnt[par ^ 1][tid] = 0;
        // ---- phase A2 (thread per env): the record, at the env's slot
        int myslot = kNoSlot;
        if (piece < kNumPieces) {
            myslot = slot_from_rank(cnt, piece, rank);
            TB_CHECK(myslot >= 0 && myslot < TILE && rank < cnt[piece]);
            uint32_t col[C];
            load_board<C, R>(sv, e, col);
            TB_CHECK(max_height_raw<C>(col) <= R);         // the record's precondition: a non-terminal board
            uint32_t *myrec = sm.rec + myslot * K::WORDS;
#pragma unroll
            for (int i = 0; i < C; ++i) myrec[K::COLX + 2 + i] = col[i];
            build_env<C, R>(s_run, myrec);
            sm.env_of[myslot] = (uint8_t)tid;
            sm.pid[myslot] = (uint8_t)piece;
            sm.vloc[myslot][0] = 0u; sm.vloc[myslot][1] = 0u;
            sm.sloc[myslot][0] = 0u; sm.sloc[myslot][1] = 0u;
        }
        if (env_thread) sm.slot_of[tid] = (uint16_t)myslot;
        __syncthreads();
        // ---- phase B (warp per window): per (piece, column loop) job; a window = EPW envs x (C - W + 1) columns
        // windows are dealt to the warps round robin, continuing across jobs (separately for jobs of 1 and of 2
        // orientations, which cost about 1 : 2), so the warps' totals differ by at most one window per class
        int woff1 = 0, woff2 = 0;
        auto column_loop = [&](auto wtag, uint32_t jb, int np, int pbase) {
            constexpr int W = decltype(wtag)::value;
            constexpr int NC = C - W + 1, EPW = 32 / NC, D = WinStride<EPW, NC, K::WORDS>::value;
            const int k = lane / NC, c = lane - k * NC;
            const bool lane_ok = k < EPW;
            const int l = (jb >> 4) & 1, n = (jb >> 8) & 3;
            const int nwin = (np + EPW - 1) / EPW;
            const int wstart = (warp + NWARPS - (n == 2 ? woff2 : woff1) % NWARPS) % NWARPS;
            if (n == 2) woff2 += nwin; else woff1 += nwin;
            for (int win = wstart; win < nwin; win += NWARPS) {
                // idle lanes (beyond EPW envs, or beyond the piece's last env) mirror the lanes of the window's env 0:
                // the same addresses as live lanes, so they add no shared-memory wavefront
                const int idx0 = window_position<EPW, D>(win, 0, np);
                const int idx = lane_ok ? window_position<EPW, D>(win, k, np) : idx0;
                const bool on = lane_ok && idx < np;
                const int slot = pbase + (on ? idx : idx0);
                TB_CHECK(idx0 >= 0 && idx0 < np && slot >= 0 && slot < TILE && pbase + np <= TILE);
                const uint32_t *rec = sm.rec + slot * K::WORDS;
                const int64_t env = e0 + (int64_t)sm.env_of[slot];
                TB_CHECK(env < sv.n_env && (int)sm.pid[slot] == (int)(jb & 15u));
                Neigh<C, R, W> nb;
                load_neigh<C, R, W>(rec, c, nb);
                // leader lane: legal / slow columns of its env, stored per orientation as the 16-bit halves of vloc / sloc
                uint16_t *const v16 = reinterpret_cast<uint16_t *>(&sm.vloc[slot][l]);
                uint16_t *const s16 = reinterpret_cast<uint16_t *>(&sm.sloc[slot][l]);
                const bool leader = on && c == 0;
#pragma unroll 1
                for (int o = 0; o < n; ++o) {
                    const int oi = (int)((jb >> 10) & 63u) + o;
                    const OriU &u = *reinterpret_cast<const OriU *>(c_odesc.w[oi]);
                    const int aslot = (int)(jb >> 16) + c * n + o;
                    bool slow = false, legal = false;
                    if (on) {
                        Eval ev;
                        const int status = eval_neigh<C, R, W>(rec, s_run, nb, u, c, ev);
                        if (status == kFastDone) { if (aslot < a_stride) emit_row<FMT>(feat_row<FMT>(feats, env, a_stride, aslot), ev, dirs); }
                        else if (status == kFastClears) slow = (!ev.terminal || want_terminal) && aslot < a_stride;
                        else slow = want_terminal && aslot < a_stride;
                        legal = !ev.terminal;
                    }
                    const uint32_t bl = __ballot_sync(FULLMASK, legal) >> (k * NC);      // cut to the env's columns by the readers
                    const uint32_t bw = __ballot_sync(FULLMASK, slow) >> (k * NC);
                    if (leader) { v16[o] = (uint16_t)bl; s16[o] = (uint16_t)bw; }
                }
            }
        };
        {
            int pbase = 0;
#pragma unroll 1
            for (int j = 0; j < kNumJobs; ++j) {
                const uint32_t jb = c_job<C>.v[j];
                const int np = cnt[jb & 15];
                if (np != 0 && ((jb >> 8) & 3u) != 0u) {
                    switch ((jb >> 5) & 7u) {
                    case 1: column_loop(IntC<1>(), jb, np, pbase); break;
                    case 2: column_loop(IntC<2>(), jb, np, pbase); break;
                    case 3: column_loop(IntC<3>(), jb, np, pbase); break;
                    default: column_loop(IntC<(C >= 4 ? 4 : 1)>(), jb, np, pbase); break;
                    }
                }
                if (j & 1) pbase += np;                    // jobs 2p, 2p + 1 = the two column loops of piece p
            }
        }
        __syncthreads();                                   // every warp is done with phase B: vloc / sloc are complete
        // ---- phase S: the placements that need the general evaluation (they clear a line), pooled over the CTA: thread t
        // counts the items of SLOT t, an inclusive scan per warp + the warps' totals give every item a global index, and
        // the items are dealt to all lanes of all warps 32 at a time -- full warps however the items are spread over the
        // slots (round 1 flattened them per warp: 13 of 32 lanes active, a fifth of K1's stall samples).
        {
            int n_active = 0;
#pragma unroll
            for (int q = 0; q < kNumPieces; ++q) n_active += cnt[q];
            uint32_t m0 = 0u, m1 = 0u;
            if (env_thread && tid < n_active) {                // slot tid: cut the marks to the loops' columns, for all readers
                const uint32_t pw = sm.piece[sm.pid[tid]];
                m0 = sm.sloc[tid][0] & loc_mask<C>(pw, 0); m1 = sm.sloc[tid][1] & loc_mask<C>(pw, 1);
                sm.sloc[tid][0] = m0; sm.sloc[tid][1] = m1;
            }
            int incl = __popc(m0) + __popc(m1);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(FULLMASK, incl, d);
                if (lane >= d) incl += v;
            }
            if (env_thread) {
                sm.sprefix[tid] = (uint16_t)incl;
                if (lane == 31) sm.wtot[warp] = incl;
            }
            __syncthreads();
            int goff[NGROUPS + 1];
            goff[0] = 0;
#pragma unroll
            for (int g = 0; g < NGROUPS; ++g) goff[g + 1] = goff[g] + sm.wtot[g];
            const int total = goff[NGROUPS];
            for (int i = tid; i < total; i += THREADS) {
                int g = 0;
#pragma unroll
                for (int q = 1; q < NGROUPS; ++q) g += goff[q] <= i;
                const int li = i - goff[g];
                const uint16_t *pre = sm.sprefix + (g << 5);
                int owner = 0;                             // first slot of the group whose inclusive prefix exceeds li
#pragma unroll
                for (int step = 16; step; step >>= 1)
                    if ((int)pre[owner + step - 1] <= li) owner += step;
                const int oslot = (g << 5) + owner;
                TB_CHECK(g < NGROUPS && owner < 32 && oslot < n_active && (int)pre[owner] > li);
                const uint32_t om0 = sm.sloc[oslot][0], om1 = sm.sloc[oslot][1];
                const int r = li - ((int)pre[owner] - __popc(om0) - __popc(om1)), p0 = __popc(om0);
                const int l = r >= p0;
                const int bit = (int)__fns(l ? om1 : om0, 0u, (l ? r - p0 : r) + 1);
                TB_CHECK(bit >= 0 && bit < 32);
                const int o = bit >> 4, cc = bit & 15;
                const uint32_t pw = sm.piece[sm.pid[oslot]];
                const int n0 = pw & 3, w0 = (pw >> 2) & 7, n1 = (pw >> 5) & 3, obase = (pw >> 10) & 63;
                const int aslot = l ? n0 * (C - w0 + 1) + cc * n1 + o : cc * n0 + o;
                Eval ev;
                // terminal afterstates (only evaluated with TB_FLAG_INCLUDE_TERMINAL) need the general wells code; without
                // the flag every item is a legal placement and the table form applies (kernel-uniform choice)
                eval_slow<C, R>(sm.rec + oslot * K::WORDS + K::COLX + 2, sm.ori[obase + (l ? n0 : 0) + o], cc, ev, nullptr,
                                want_terminal ? nullptr : s_run);
                TB_CHECK(aslot < a_stride && e0 + (int64_t)sm.env_of[oslot] < sv.n_env);
                emit_row<FMT>(feat_row<FMT>(feats, e0 + (int64_t)sm.env_of[oslot], a_stride, aslot), ev, dirs);
            }
        }
        // ---- legal-action masks, thread per env (coalesced)
        if (env_thread && e < sv.n_env) {
            unsigned long long v = 0ull;
            if (myslot != kNoSlot) {
                const uint32_t pw = sm.piece[piece];
                const int n0 = (int)(pw & 3u), n1 = (int)((pw >> 5) & 3u);
                const int s1 = n0 * (C - (int)((pw >> 2) & 7u) + 1);
                v = (unsigned long long)slots_of(sm.vloc[myslot][0] & loc_mask<C>(pw, 0), n0) |
                    ((unsigned long long)slots_of(sm.vloc[myslot][1] & loc_mask<C>(pw, 1), n1) << s1);
            }
            if (valid_out) valid_out[e] = v;
            if (count_out) count_out[e] = __popcll(v);
        }
        // the next tile's phase A2 overwrites the records: it starts behind the barrier after phase A1
    }
}

template <int C, int R, int FMT, int TILE, int MINB, int THREADS = TILE>
__global__ void __launch_bounds__(THREADS, MINB)
k_afterstates(StateView sv, float *__restrict__ feats, unsigned long long *__restrict__ valid_out,
              int *__restrict__ count_out, int a_stride, F8 dirs, int flags)
{
    afterstates_body<C, R, FMT, TILE, THREADS>(sv, feats, valid_out, count_out, a_stride, dirs, flags);
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
    if (mt.piece >= kNumPieces) return;                 // finished / void fork: no afterstates
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
// K2 step.
//
// A CTA owns a tile of TILE envs, thread per env, in two passes of the same shape:
//   pass 0  load + transpose the board; envs whose stack is within 4 rows of the top ("tall": a placement may be
//           terminal) are appended to a list, their columns staged in shared memory
//   pooled  the legality masks of the CTA's tall envs, one thread per LISTED env: the warps that hold list entries run
//           the column-parallel legality test with full lanes instead of every warp running it for its 2-3 tall lanes
//           (round 1: 39 % of K2's warp-instructions ran with 3 of 32 lanes), the others wait at the barrier
//   pass 1  rank-select the action among the legal slots (game.py:83), place + clear, the chosen afterstate's features
//           from scratch (cumulative wells by table: no data-dependent POPC loop), next piece; envs that are tall NOW
//           are listed again
//   pooled  same code (one copy: the pass loop is rolled), now deciding game over (game.py:94-100)
//   finish  reward / done, auto-reset, stores.
// On a low board every slot is legal and neither pooled phase has work for the env.
// ---------------------------------------------------------------------------------------------
template <int C, int R, int TILE>
struct StepSmem {
    uint32_t cols[TILE][C | 1];          // columns of the listed envs (odd stride)
    alignas(16) uint32_t run_raw[2 * RunTab<R>::SIZE];   // the plain table (no permuted copy: a CTA here handles one tile
                                                         // and staging 8 KB more costs 5 %), placed at a multiple of its size
    unsigned long long vmask[TILE];      // by env: legal slots, written by the pooled phase
    uint16_t list[TILE];                 // listed (tall) envs of the current pass
    uint8_t lpiece[TILE];                // by env: piece to test in the pooled phase
    int n_list[2];
    uint32_t ori[32], piece[16];
};

// named barriers (ids 1..15; 0 is __syncthreads): arrive = signal without waiting
template <int C, int R, int TILE, int MINB>
__global__ void __launch_bounds__(TILE, MINB)
k_step(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, const int32_t *__restrict__ actions,
       const uint8_t *__restrict__ tape, float *__restrict__ obs, int32_t *__restrict__ reward,
       uint8_t *__restrict__ done, int32_t *__restrict__ lines, int32_t *status, int flags, F8 dirs)
{
    __shared__ StepSmem<C, R, TILE> sm;
    const int tid = threadIdx.x;
    if (tid < kNumOris) sm.ori[tid] = c_ori[tid];
    if (tid < kNumPieces) sm.piece[tid] = c_piece[tid];
    if (tid < 2) sm.n_list[tid] = 0;
    uint32_t *const s_run = run_tab_place<R>(sm.run_raw);
    stage_run_table<R, false>(s_run);
    __syncthreads();
    const int64_t e = (int64_t)blockIdx.x * TILE + tid;
    const bool in_range = e < sv.n_env;
    uint32_t col[C];
    Meta mt; mt.piece = kPieceVoid; mt.bag = 0u; mt.draws = 0u;
    uint2 ep = make_uint2(0u, 0u);
    bool listed = false, go = false;
    int lc = 0;

#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        listed = false;
        if (pass == 0) {
            if (in_range) {
                load_board<C, R>(sv, e, col);
                mt = unpack_meta(sv.meta[e]);
                ep = sv.epi[e];
                listed = mt.piece < kNumPieces && max_height<C>(col) + 4 > R;
            }
        } else {
            // ---- the action (game.py:83), against the legal slots of the env's piece
            const bool live = mt.piece < kNumPieces;                          // not a finished rollout fork
            const uint32_t pw = sm.piece[live ? mt.piece : 0];
            unsigned long long vm = 0ull;
            if (live) vm = max_height<C>(col) + 4 > R ? sm.vmask[tid] : (1ull << piece_num_slots(pw, C)) - 1ull;
            int sel = -1;
            if (in_range) {
                const int action = actions[e];
                if (action >= 0) {
                    if (flags & TB_FLAG_ACTION_IS_SLOT) {
                        if (action < 64 && ((vm >> action) & 1ull)) sel = action;
                    } else if (action < __popcll(vm)) {
                        sel = nth_set_bit(vm, action);
                    }
                }
                if (sel < 0) {
                    // IndexError in the reference (game.py:83).  The env is left untouched, its outputs are defined (zero
                    // observation / reward / lines; done = it has no legal placement at all), and status reports the
                    // lowest offending env as 0x7FFFFFFF - env.
                    if (status) atomicMax(status, 0x7FFFFFFF - (int)(e < 0x7FFFFFFE ? e : 0x7FFFFFFE));
                    if (!(flags & TB_FLAG_VALIDATE_ONLY)) {
                        if (obs) {
                            float4 *o = reinterpret_cast<float4 *>(obs + e * 8);
                            o[0] = make_float4(0.f, 0.f, 0.f, 0.f); o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        if (reward) reward[e] = 0;
                        if (done) done[e] = (uint8_t)(vm == 0ull);
                        if (lines) lines[e] = 0;
                    }
                }
            }
            if (flags & TB_FLAG_VALIDATE_ONLY) return;                        // dry run (kernel-uniform): status only
            go = sel >= 0;
            if (go) {
                int ori, c;
                slot_to_placement(pw, C, sel, ori, c);
                Eval ev;
                eval_slow<C, R, false>(col, sm.ori[ori], c, ev, col, s_run);       // current_state = afterstates[action]
                lc = popc32(ev.full);                                         // game.py:85
                if (obs) {                                                    // game.py:91 (stored here: fewer live registers)
                    float4 *o = reinterpret_cast<float4 *>(obs + e * 8);
                    o[0] = make_float4(ev.f[0] * dirs.v[0], ev.f[1] * dirs.v[1], ev.f[2] * dirs.v[2], ev.f[3] * dirs.v[3]);
                    o[1] = make_float4(ev.f[4] * dirs.v[4], ev.f[5] * dirs.v[5], ev.f[6] * dirs.v[6], ev.f[7] * dirs.v[7]);
                }
                if (lines) lines[e] = lc;
                mt.piece = draw_piece(piece_set, env_key(seed, (uint64_t)(env_offset + e)), mt, tape, e);   // game.py:87
                listed = mt.piece < kNumPieces && max_height<C>(col) + 4 > R; // else: every slot of the new piece is legal
            }
        }
        if (listed) {
#pragma unroll
            for (int k = 0; k < C; ++k) sm.cols[tid][k] = col[k];
            sm.lpiece[tid] = (uint8_t)mt.piece;
            const int li = atomicAdd(&sm.n_list[pass], 1);
            TB_CHECK(li >= 0 && li < TILE);
            sm.list[li] = (uint16_t)tid;
        }
        // the barrier that completes the list also tells whether the tile listed anything: on low boards (greedy play)
        // most tiles list nothing and skip the pooled phase together with its second barrier
        __syncthreads();
        if (sm.n_list[pass] == 0) continue;
        // ---- pooled: legal slots of the listed envs (game.py:69 / :94-100), one thread per list entry.  (Measured and
        // rejected, profiles/README.md r2f / r2l: a warp per entry with lane = enumeration slot -- shorter wait at the
        // barrier, more instructions in total, 0.148 vs 0.122 ms; split barriers where only the consumer warps and the
        // warps that listed an env wait (bar.arrive for the rest) -- 0.089 vs 0.085 ms.)
        if (tid < sm.n_list[pass]) {
            const int env = (int)sm.list[tid];
            TB_CHECK(env >= 0 && env < TILE && sm.lpiece[env] < kNumPieces);
            uint32_t c2[C];
#pragma unroll
            for (int k = 0; k < C; ++k) c2[k] = sm.cols[env][k];
            sm.vmask[env] = valid_slots<C, R>(c2, sm.piece[sm.lpiece[env]], sm.ori);
        }
        __syncthreads();
    }
    if (!go) return;
    // ---- finish: game over (game.py:88), reward, auto-reset (example_play.py:20-21), stores
    const bool dn = mt.piece >= kNumPieces || (listed && sm.vmask[tid] == 0ull);
    int rew = lc - 1 - (dn ? 100 : 0);                                        // game.py:86,89-90
    ep.x += 1u; ep.y += (uint32_t)lc;
    if (reward) reward[e] = rew;
    if (done) done[e] = (uint8_t)dn;
    if (dn && (flags & TB_FLAG_AUTO_RESET) && !tape) {
#pragma unroll
        for (int k = 0; k < C; ++k) col[k] = 0u;
        mt.piece = draw_piece(piece_set, env_key(seed, (uint64_t)(env_offset + e)), mt, nullptr, e);
        ep = make_uint2(0u, 0u);
    }
    store_board<C, R>(sv, e, col);
    sv.meta[e] = pack_meta<C>(col, mt);
    sv.epi[e] = ep;
}

// K2, thread per env without CTA-level cooperation (round 1's structure, plus the table-based evaluation of the chosen
// afterstate): no barriers after the table staging, two inlined copies of the legality test.  Kept as tuning
// configuration k2_cfg = 5 for A/B runs.
template <int C, int R>
__global__ void __launch_bounds__(128, 8)
k_step_tpe(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, const int32_t *__restrict__ actions,
       const uint8_t *__restrict__ tape, float *__restrict__ obs, int32_t *__restrict__ reward,
       uint8_t *__restrict__ done, int32_t *__restrict__ lines, int32_t *status, int flags, F8 dirs)
{
    __shared__ uint32_t s_ori[32], s_piece[16];
    __shared__ __align__(16) uint32_t s_run_raw[2 * RunTab<R>::SIZE];
    uint32_t *const s_run = run_tab_place<R>(s_run_raw);
    stage_run_table<R, false>(s_run);
    stage_tables(s_ori, s_piece);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= sv.n_env) return;
    uint32_t col[C];
    load_board<C, R>(sv, e, col);
    Meta mt = unpack_meta(sv.meta[e]);
    uint2 ep = sv.epi[e];
    const bool live = mt.piece < kNumPieces;                                  // not a finished rollout fork
    const uint32_t pw = s_piece[live ? mt.piece : 0];
    const int action = actions[e];
    const unsigned long long vm = live ? valid_mask<C, R>(col, pw, s_ori) : 0ull;   // game.py:69
    int sel = -1;
    if (action >= 0) {
        if (flags & TB_FLAG_ACTION_IS_SLOT) {
            if (action < 64 && ((vm >> action) & 1ull)) sel = action;
        } else if (action < __popcll(vm)) {
            sel = nth_set_bit(vm, action);                                    // game.py:83
        }
    }
    if (sel < 0) {
        // IndexError in the reference (game.py:83).  The env is left untouched, its outputs are defined (zero
        // observation / reward / lines; done = it has no legal placement at all), and status reports the lowest
        // offending env as 0x7FFFFFFF - env.
        if (status) atomicMax(status, 0x7FFFFFFF - (int)(e < 0x7FFFFFFE ? e : 0x7FFFFFFE));
        if (!(flags & TB_FLAG_VALIDATE_ONLY)) {
            if (obs) {
                float4 *o = reinterpret_cast<float4 *>(obs + e * 8);
                o[0] = make_float4(0.f, 0.f, 0.f, 0.f); o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (reward) reward[e] = 0;
            if (done) done[e] = (uint8_t)(vm == 0ull);
            if (lines) lines[e] = 0;
        }
        return;
    }
    if (flags & TB_FLAG_VALIDATE_ONLY) return;                                // dry run: only the status is produced
    int ori, c;
    slot_to_placement(pw, C, sel, ori, c);
    Eval ev;
    eval_slow<C, R, false>(col, s_ori[ori], c, ev, col, s_run);                             // current_state = afterstates[action]
    const int lc = popc32(ev.full);                                           // game.py:85
    int rew = lc - 1;                                                         // game.py:86
    const uint64_t key = env_key(seed, (uint64_t)(env_offset + e));
    mt.piece = draw_piece(piece_set, key, mt, tape, e);                       // game.py:87
    const bool dn = mt.piece >= kNumPieces || !any_valid<C, R>(col, s_piece[mt.piece], s_ori);   // game.py:88,94-100
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

// Episode statistics of one step, aggregated over the warp into its shared-memory vector (ballots are warp-uniform; lane 0
// accumulates).  Must be reached by the whole warp.  placed: the lane's env made a placement; dn: it ended its episode
// with it (ep_done = the finished episode's placements / lines); lc: lines cleared; n_slots: afterstates enumerated.
__device__ __forceinline__ void warp_step_stats(long long *wstat, int lane, bool placed, bool dn, int lc, int n_slots,
                                                uint2 ep_done)
{
    const unsigned bp = __ballot_sync(FULLMASK, placed), bd = __ballot_sync(FULLMASK, dn);
    const unsigned b1 = __ballot_sync(FULLMASK, lc == 1), b2 = __ballot_sync(FULLMASK, lc == 2);
    const unsigned b3 = __ballot_sync(FULLMASK, lc == 3), b4 = __ballot_sync(FULLMASK, lc == 4);
    const int n_after = __reduce_add_sync(FULLMASK, n_slots);
    long long se = 0, sl = 0, me = 0, ml = 0;
    if (bd) {
        se = warp_sum(dn ? (long long)ep_done.x : 0ll); sl = warp_sum(dn ? (long long)ep_done.y : 0ll);
        me = warp_max(dn ? (long long)ep_done.x : 0ll); ml = warp_max(dn ? (long long)ep_done.y : 0ll);
    }
    if (lane == 0) {
        const int np_ = __popc(bp), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3), n4 = __popc(b4);
        const int nd = __popc(bd), nl = n1 + 2 * n2 + 3 * n3 + 4 * n4;
        wstat[TB_ST_PLACEMENTS] += np_; wstat[TB_ST_EPISODES] += nd; wstat[TB_ST_LINES] += nl;
        wstat[TB_ST_REWARD] += nl - np_ - 100 * nd; wstat[TB_ST_AFTERSTATES] += n_after;
        wstat[TB_ST_LINES0] += np_ - n1 - n2 - n3 - n4; wstat[TB_ST_LINES1] += n1; wstat[TB_ST_LINES2] += n2;
        wstat[TB_ST_LINES3] += n3; wstat[TB_ST_LINES4] += n4;
        if (bd) {
            wstat[TB_ST_SUM_EP_STEPS] += se; wstat[TB_ST_SUM_EP_LINES] += sl;
            if (me > wstat[TB_ST_MAX_EP_STEPS]) wstat[TB_ST_MAX_EP_STEPS] = me;
            if (ml > wstat[TB_ST_MAX_EP_LINES]) wstat[TB_ST_MAX_EP_LINES] = ml;
        }
    }
}
// a warp's statistics vector -> global (sums, the two maxima by max)
__device__ __forceinline__ void warp_stats_to_global(const long long *wstat, int lane, int64_t *stats)
{
    __syncwarp();
    if (lane < TB_ST_COUNT) {
        const long long r = wstat[lane];
        if (r != 0) {
            if (lane == TB_ST_MAX_EP_LINES || lane == TB_ST_MAX_EP_STEPS) atomicMax((long long *)&stats[lane], r);
            else atomicAdd((unsigned long long *)&stats[lane], (unsigned long long)r);
        }
    }
}

// Apply the chosen placement to the lane's env: lock, clear, reward, next piece, game-over, auto-reset.
// Returns the legal-slot mask of the NEW piece on the NEW board (never 0: a finished env is reset in place).
template <int C, int R>
__device__ __forceinline__ unsigned long long
apply_placement(uint32_t *col, Meta &mt, uint2 &ep, uint32_t d, int c, int piece_set, uint64_t key,
                const uint32_t *s_ori, const uint32_t *s_piece, LaneStats &st, bool no_reset = false,
                const uint8_t *tape_piece = nullptr)
{
    int a, term;
    uint32_t full;
    place_and_clear<C, R>(col, d, c, a, full, term);
    const int lc = popc32(full);
    int rew = lc - 1;
    mt.piece = draw_piece(piece_set, key, mt, tape_piece, 0);
    unsigned long long vm = mt.piece < kNumPieces ? valid_mask<C, R>(col, s_piece[mt.piece], s_ori) : 0ull;
    const bool dn = vm == 0ull;
    if (dn) rew -= 100;
    ep.x += 1u; ep.y += (uint32_t)lc;
    st.placements += 1; st.lines += lc; st.reward += rew;
    st.l0 += (lc == 0); st.l1 += (lc == 1); st.l2 += (lc == 2); st.l3 += (lc == 3); st.l4 += (lc == 4);
    if (dn) {
        st.episodes += 1;
        st.sum_ep_steps += ep.x; st.sum_ep_lines += ep.y;
        st.max_ep_lines = imax(st.max_ep_lines, (int)ep.y);
        st.max_ep_steps = imax(st.max_ep_steps, (int)ep.x);
        if (no_reset) { mt.piece = kPieceDead; return 0ull; }             // the env stays finished (rollout forks)
#pragma unroll
        for (int k = 0; k < C; ++k) col[k] = 0u;
        mt.piece = set_piece(piece_set, bag_draw(set_size(piece_set), key, mt.bag, mt.draws));
        ep = make_uint2(0u, 0u);
        vm = (1ull << piece_num_slots(s_piece[mt.piece], C)) - 1ull;      // empty board: everything is legal
    }
    return vm;
}

// (Round 2 tried two CTA-cooperative schemes here, both bit-exact and both slower than this kernel's 2.56 ms per 32 steps of
// 2^20 envs: the K2 scheme -- a CTA steps a tile in lockstep and pools the legality test of its tall envs, two barriers
// per step: 3.31 ms -- and re-packing the tile's envs over the lanes by height class every few steps so that warps are
// all-low or all-tall: 3.45 ms re-packing every 4 steps, 2.76 ms never re-packing (256-thread CTAs).  40 % of the
// warp-instructions here run with 5 of 32 lanes (the tall-board legality test), but warps progress at very different
// rates and every CTA-level barrier makes all of them wait for the slowest: independent warps with 6 CTAs per SM hide the
// divergence better than cooperation removes it.  profiles/README.md, r2d / r2k.)
// random policy: everything is per-env, one thread per env, board in registers for all n_steps
template <int C, int R>
__global__ void __launch_bounds__(128, 6)                  // 80 registers, no spills (measured: profiles/README.md, r1g)
k_rollout_random(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, int n_steps, int64_t *stats,
                 int no_reset, const uint8_t *__restrict__ tape, int tape_stride)
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
        if (mt.piece >= kNumPieces) continue;              // finished / void fork: untouched
        uint2 ep = sv.epi[e];
        const uint64_t key = env_key(seed, (uint64_t)(env_offset + e));
        unsigned long long vm = valid_mask<C, R>(col, s_piece[mt.piece], s_ori);
        for (int t = 0; t < n_steps; ++t) {
            if (mt.piece >= kNumPieces) break;
            const uint32_t pw = s_piece[mt.piece];
            st.afterstates += piece_num_slots(pw, C);
            if (vm == 0ull) {
                if (no_reset) { mt.piece = kPieceDead; break; }
                // no legal placement: only reachable from a caller-supplied dead state -> start a new episode
#pragma unroll
                for (int k = 0; k < C; ++k) col[k] = 0u;
                mt.piece = set_piece(piece_set, bag_draw(set_size(piece_set), key, mt.bag, mt.draws));
                ep = make_uint2(0u, 0u);
                vm = (1ull << piece_num_slots(s_piece[mt.piece], C)) - 1ull;
                continue;
            }
            const int action = (int)bounded(rng32(key, mt.draws, 1u), (uint32_t)__popcll(vm));
            const int slot = nth_set_bit(vm, action);
            int ori, c;
            slot_to_placement(pw, C, slot, ori, c);
            vm = apply_placement<C, R>(col, mt, ep, s_ori[ori], c, piece_set, key, s_ori, s_piece, st, no_reset != 0,
                                       tape ? tape + e * tape_stride + t : nullptr);
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
__device__ __forceinline__ unsigned long long score_key(uint32_t ord, int slot)
{
    return ((unsigned long long)ord << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)slot);   // ties: lower slot
}

// Greedy linear policy.  Same tile scheme as K1; thread tid owns env tid of the tile for all n_steps.  Per step:
//   A (thread per env)  counting sort by piece (rank by shared atomic, barrier, slot), record built at the env's slot
//   B (warp per window of same-piece envs)  score every legal placement, keep the first arg-max per env and column loop
//   S (per warp, its own 32 envs)  the line-clearing placements, from scratch
//   C (thread per env)  apply the chosen placement, next piece, game over / auto-reset; the board stays in registers
//                       until phase A of the next step writes it into the env's NEW slot
// Episode statistics are aggregated per warp in shared memory.  THREADS > TILE: small-batch configuration, see K1.
template <int C, int R, int TILE, int MINB, int THREADS = TILE>
__global__ void __launch_bounds__(THREADS, MINB)
k_rollout_greedy(StateView sv, int64_t env_offset, uint64_t seed, int piece_set, int n_steps, F8 wts, int64_t *stats,
                 int no_reset, const uint8_t *__restrict__ tape, int tape_stride)
{
    using K = Env<C, R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(TILE % 32 == 0 && THREADS % TILE == 0, "TILE: whole warps of envs; THREADS: a multiple of TILE");
    constexpr int NWARPS = THREADS / 32, NGROUPS = TILE / 32, WPG = NWARPS / NGROUPS;
    CtaSmem<C, R, TILE> &sm = *reinterpret_cast<CtaSmem<C, R, TILE> *>(smem_raw);
    BestSmem<TILE, THREADS> &bs =
        *reinterpret_cast<BestSmem<TILE, THREADS> *>(smem_raw + ((sizeof(CtaSmem<C, R, TILE>) + 15) & ~(size_t)15));
    __shared__ __align__(16) uint32_t s_run_raw[RunTab<R>::SIZE + RunTab<R>::WORDS];
    uint32_t *const s_run = run_tab_place<R>(s_run_raw);
    stage_cta(sm, s_run);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool env_thread = THREADS == TILE || tid < TILE;   // this thread owns env `tid` of the tile in the per-env phases
    const int64_t n_tiles = (sv.n_env + TILE - 1) / TILE;
    long long *wstat = bs.wstat[warp];
    if (lane < TB_ST_COUNT) wstat[lane] = 0;
    int par = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e = tile * TILE + tid;
        const bool in_range = env_thread && e < sv.n_env;
        bool active = in_range;                            // false once finished in a no-reset rollout
        Meta mt; mt.piece = 0; mt.bag = 0u; mt.draws = 0u;
        uint2 ep = make_uint2(0u, 0u);
        uint32_t col[C];                                   // the env's board between phase C and the next phase A
        if (in_range) {
            load_board<C, R>(sv, e, col);
            mt = unpack_meta(sv.meta[e]);
            ep = sv.epi[e];
            active = mt.piece < kNumPieces;
        } else {
#pragma unroll
            for (int i = 0; i < C; ++i) col[i] = 0u;
        }
        for (int t = 0; t < n_steps; ++t, par ^= 1) {
            int *cnt = sm.cnt[par];
            // ---- phase A1: rank among the tile's envs with the same piece.  cnt[par] was zeroed during the previous A2.
            int rank = 0;
            if (active) rank = atomicAdd(&cnt[mt.piece], 1);
            __syncthreads();                               // counts complete; nobody reads the previous step's records
            if (tid <= kNumPieces) sm.cnt[par ^ 1][tid] = 0;
            // ---- phase A2: the record, at the env's slot
            int myslot = kNoSlot, n_slots = 0;
            if (active) {
                myslot = slot_from_rank(cnt, mt.piece, rank);
                TB_CHECK(myslot >= 0 && myslot < TILE && rank < cnt[mt.piece]);
                TB_CHECK(max_height_raw<C>(col) <= R);     // the record's precondition: a non-terminal board
                uint32_t *myrec = sm.rec + myslot * K::WORDS;
#pragma unroll
                for (int i = 0; i < C; ++i) myrec[K::COLX + 2 + i] = col[i];
                build_env<C, R>(s_run, myrec);
                sm.pid[myslot] = (uint8_t)mt.piece;
                sm.vloc[myslot][0] = 0u; sm.vloc[myslot][1] = 0u;
                sm.sloc[myslot][0] = 0u; sm.sloc[myslot][1] = 0u;
                bs.best[myslot] = 0ull;
                n_slots = piece_num_slots(sm.piece[mt.piece], C);
            }
            if (WPG > 1 && env_thread) sm.slot_of[tid] = (uint16_t)myslot;
            __syncthreads();
            // ---- phase B: score every legal placement, keep the first arg-max per env and column loop
            int woff1 = 0, woff2 = 0;                      // round robin continued across jobs, as in K1
            auto column_loop = [&](auto wtag, uint32_t jb, int np, int pbase) {
                constexpr int W = decltype(wtag)::value;
                constexpr int NC = C - W + 1, EPW = 32 / NC, D = WinStride<EPW, NC, K::WORDS>::value;
                const int k = lane / NC, c = lane - k * NC;
                const bool lane_ok = k < EPW;
                const int l = (jb >> 4) & 1, n = (jb >> 8) & 3;
                const int nwin = (np + EPW - 1) / EPW;
                const int wstart = (warp + NWARPS - (n == 2 ? woff2 : woff1) % NWARPS) % NWARPS;
                if (n == 2) woff2 += nwin; else woff1 += nwin;
                for (int win = wstart; win < nwin; win += NWARPS) {
                    // idle lanes mirror the lanes of the window's env 0 (same addresses: no extra wavefront), see K1
                    const int idx0 = window_position<EPW, D>(win, 0, np);
                    const int idx = lane_ok ? window_position<EPW, D>(win, k, np) : idx0;
                    const bool on = lane_ok && idx < np;
                    const int slot = pbase + (on ? idx : idx0);
                    TB_CHECK(idx0 >= 0 && idx0 < np && slot >= 0 && slot < TILE && pbase + np <= TILE);
                    TB_CHECK((int)sm.pid[slot] == (int)(jb & 15u));
                    const uint32_t *rec = sm.rec + slot * K::WORDS;
                    Neigh<C, R, W> nb;
                    load_neigh<C, R, W>(rec, c, nb);
                    uint32_t best_ord = 0u;                // this lane's best orderable score (0 = none) and its slot
                    int best_slot = 0;
                    uint16_t *const s16 = reinterpret_cast<uint16_t *>(&sm.sloc[slot][l]);
                    const bool leader = on && c == 0;
#pragma unroll 1
                    for (int o = 0; o < n; ++o) {
                        const int oi = (int)((jb >> 10) & 63u) + o;
                        const OriU &u = *reinterpret_cast<const OriU *>(sm.odesc[oi]);
                        const int aslot = (int)(jb >> 16) + c * n + o;
                        bool slow = false;
                        if (on) {
                            Eval ev;
                            const int status = eval_neigh<C, R, W>(rec, s_run, nb, u, c, ev);
                            if (status == kFastDone) {
                                const uint32_t ord = orderable(fitness(ev.f, wts.v));         // game.py:109-120
                                if (ord > best_ord) { best_ord = ord; best_slot = aslot; }    // slots ascend with o
                            } else if (status == kFastClears) {
                                slow = !ev.terminal;
                            }
                        }
                        const uint32_t bw = __ballot_sync(FULLMASK, slow) >> (k * NC);   // cut to the env's columns by phase S
                        if (leader) s16[o] = (uint16_t)bw;
                    }
                    // first arg-max over the env's NC lanes: highest score, then lowest slot (= lowest column, then o)
                    uint32_t m = best_ord;
#pragma unroll
                    for (int d = 1; d < NC; d <<= 1) {
                        const uint32_t m2 = __shfl_down_sync(FULLMASK, m, d);
                        if (c + d < NC) m = max(m, m2);
                    }
                    m = __shfl_sync(FULLMASK, m, k * NC);                      // the segment's maximum, from its leader
                    const uint32_t hit = (__ballot_sync(FULLMASK, best_ord == m) >> (k * NC)) & ((1u << NC) - 1u);
                    const int src = k * NC + __ffs((int)hit) - 1;              // lowest column that reaches it
                    const int sl = __shfl_sync(FULLMASK, best_slot, src & 31);
                    if (leader) { sm.vloc[slot][l] = m; bs.bslot[slot][l] = (uint8_t)sl; }
                }
            };
            {
                int pbase = 0;
#pragma unroll 1
                for (int j = 0; j < kNumJobs; ++j) {
                    const uint32_t jb = sm.job[j];
                    const int np = cnt[jb & 15];
                    if (np != 0 && ((jb >> 8) & 3u) != 0u) {
                        switch ((jb >> 5) & 7u) {
                        case 1: column_loop(IntC<1>(), jb, np, pbase); break;
                        case 2: column_loop(IntC<2>(), jb, np, pbase); break;
                        case 3: column_loop(IntC<3>(), jb, np, pbase); break;
                        default: column_loop(IntC<(C >= 4 ? 4 : 1)>(), jb, np, pbase); break;
                        }
                    }
                    if (j & 1) pbase += np;
                }
            }
            __syncthreads();                               // phase B is complete: vloc / bslot / sloc of every slot
            // ---- phase S: the line-clearing legal placements of this warp's own 32 envs, one lane per item.  Their best
            // keys go to best[slot of the env]: only this warp (this group's warps) touches those entries.  (Pooling the
            // items over the CTA as K1 does -- in the greedy steady state this phase is 9 % of K3's instructions at 14 of
            // 32 lanes -- costs two more barriers per step and measured -0.3 %: profiles/README.md, r2v.)
            const int grp = WPG == 1 ? warp : warp % NGROUPS, sub = WPG == 1 ? 0 : warp / NGROUPS;
            const int gslot = WPG == 1 ? myslot : (int)sm.slot_of[(grp << 5) + lane];
            const bool has = gslot != kNoSlot;
            const uint32_t gpw = has ? sm.piece[sm.pid[gslot]] : 0u;
            for_each_slow_item(has ? sm.sloc[gslot][0] & loc_mask<C>(gpw, 0) : 0u, has ? sm.sloc[gslot][1] & loc_mask<C>(gpw, 1) : 0u,
                               gslot, lane, sub, WPG,
                               [&](int owner, int oslot, int l, int o, int cc) {
                const uint32_t pw = sm.piece[sm.pid[oslot]];
                const int n0 = pw & 3, w0 = (pw >> 2) & 7, n1 = (pw >> 5) & 3, obase = (pw >> 10) & 63;
                const int aslot = l ? n0 * (C - w0 + 1) + cc * n1 + o : cc * n0 + o;
                Eval ev;
                eval_slow<C, R>(sm.rec + oslot * K::WORDS + K::COLX + 2, sm.ori[obase + (l ? n0 : 0) + o], cc, ev, nullptr,
                                s_run);                   // legal placements only: non-terminal, wells by table
                atomicMax(&bs.best[oslot], score_key(orderable(fitness(ev.f, wts.v)), aslot));
            });
            if (WPG == 1) __syncwarp();                    // phase C reads the best keys of this warp's envs only
            else __syncthreads();                          // ... or of a group that several warps worked on
            // ---- phase C (thread per env): apply the choice, draw the next piece, game over / auto-reset
            int lc = 0;
            bool placed = false, dn = false;
            uint2 ep_done = make_uint2(0u, 0u);
            if (active) {
                // the board comes back from the record (not kept in registers across phase B)
                const uint32_t *myrec = sm.rec + myslot * K::WORDS;
#pragma unroll
                for (int i = 0; i < C; ++i) col[i] = myrec[K::COLX + 2 + i];
                const uint64_t key = env_key(seed, (uint64_t)(env_offset + e));
                unsigned long long best = bs.best[myslot];
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const uint32_t m = sm.vloc[myslot][l];
                    if (m != 0u) {
                        const unsigned long long k64 = score_key(m, (int)bs.bslot[myslot][l]);
                        best = k64 > best ? k64 : best;
                    }
                }
                if (best != 0ull) {
                    const int aslot = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
                    int ori, cc, a, term;
                    uint32_t full;
                    slot_to_placement(sm.piece[mt.piece], C, aslot, ori, cc);
                    place_and_clear<C, R>(col, sm.ori[ori], cc, a, full, term);
                    lc = popc32(full);
                    placed = true;
                    mt.piece = draw_piece(piece_set, key, mt, tape ? tape + e * tape_stride + t : nullptr, 0);
                    dn = mt.piece >= kNumPieces || !any_valid<C, R>(col, sm.piece[mt.piece], sm.ori);
                    ep.x += 1u; ep.y += (uint32_t)lc;
                    ep_done = ep;
                }
                if ((dn || !placed) && no_reset) {
                    mt.piece = kPieceDead;                  // stays finished (rollout forks)
                    active = false;
                } else if (dn || !placed) {
                    // game over (or a caller-supplied dead state): start a new episode in place
#pragma unroll
                    for (int i = 0; i < C; ++i) col[i] = 0u;
                    mt.piece = set_piece(piece_set, bag_draw(set_size(piece_set), key, mt.bag, mt.draws));
                    ep = make_uint2(0u, 0u);
                }
            }
            warp_step_stats(wstat, lane, placed, dn, lc, n_slots, ep_done);
        }
        if (in_range && (active || mt.piece == kPieceDead)) {
            store_board<C, R>(sv, e, col);
            sv.meta[e] = pack_meta<C>(col, mt);
            sv.epi[e] = ep;
        }
    }
    warp_stats_to_global(wstat, lane, stats);
}

// ---------------------------------------------------------------------------------------------
// Rollout forks (Tetris.single_rollout / perform_rollouts, game.py:129-160, for every env and every action).
// Child d = (env * a_stride + slot) * n_forks + fork starts from the parent's board with `slot` applied, the
// parent's bag, and its own RNG stream (seed2, child id); the next piece is drawn and tested for game over.
// ---------------------------------------------------------------------------------------------
template <int C, int R>
__global__ void __launch_bounds__(128)
k_fork(StateView parent, StateView child, int a_stride, int n_forks, uint64_t seed2, int64_t child_offset, int piece_set,
       const uint8_t *__restrict__ tape, int tape_stride)
{
    __shared__ uint32_t s_ori[32], s_piece[16];
    stage_tables(s_ori, s_piece);
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= child.n_env) return;
    const int64_t e = d / ((int64_t)a_stride * n_forks);
    const int slot = (int)((d / n_forks) % a_stride);
    uint32_t col[C];
    load_board<C, R>(parent, e, col);
    Meta mt = unpack_meta(parent.meta[e]);
    bool ok = mt.piece < kNumPieces;
    int ori = 0, c = 0;
    if (ok) {
        const uint32_t pw = s_piece[mt.piece];
        ok = slot < piece_num_slots(pw, C);
        if (ok) {
            slot_to_placement(pw, C, slot, ori, c);
            ok = placement_valid<C, R>(col, s_ori[ori], c, max_height<C>(col));
        }
    }
    if (!ok) {
#pragma unroll
        for (int k = 0; k < C; ++k) col[k] = 0u;
        mt.piece = kPieceVoid;
    } else {
        int a, term;
        uint32_t full;
        place_and_clear<C, R>(col, s_ori[ori], c, a, full, term);          // self.step(action): game.py:83
        const uint64_t key = env_key(seed2, (uint64_t)(child_offset + d));
        mt.piece = draw_piece(piece_set, key, mt, tape ? tape + d * tape_stride : nullptr, 0);   // :87
        if (mt.piece < kNumPieces && !any_valid<C, R>(col, s_piece[mt.piece], s_ori)) mt.piece = kPieceDead;   // :88, :133-137
    }
    store_board<C, R>(child, d, col);
    child.meta[d] = pack_meta<C>(col, mt);
    child.epi[d] = make_uint2(0u, 0u);
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
    if (rows_out || heights_out) {
        uint32_t w[S::NW];
#pragma unroll
        for (int b = 0; b < S::NB; ++b) {
            const uint4 v = sv.planes[(int64_t)b * sv.n_env + e];
            w[4 * b] = v.x; w[4 * b + 1] = v.y; w[4 * b + 2] = v.z; w[4 * b + 3] = v.w;
        }
        if (rows_out) {
#pragma unroll
            for (int r = 0; r < S::N; ++r) rows_out[i * S::N + r] = (uint16_t)(w[r >> 1] >> (16 * (r & 1)));
        }
        if (heights_out) {                                   // lowest_free_rows (state.py:162-172), from the board itself
            uint32_t col[C];
            rows_to_cols<C, R>(w, col);
#pragma unroll
            for (int c = 0; c < C; ++c) heights_out[i * C + c] = (uint8_t)height_of(col[c]);
        }
    }
    if (piece_out) piece_out[i] = (uint8_t)((sv.meta[e].z >> 16) & 0xffu);
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
    if (piece_in) mt.piece = checked_piece((int)piece_in[i]);
    // a board with a cell at or above row R is a terminal state (state.py:111-117): nothing is placed on it any more
    // (every kernel relies on column heights <= R)
    if (max_height<C>(col) > R && mt.piece < kNumPieces) mt.piece = kPieceDead;
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


// =============================================================================================
// Launchers of one board shape (the entries of its TbShapeVT).  Argument checks that do not depend on the shape are
// done by the ABI layer (tb_abi.cu) before these are reached.
// =============================================================================================
static inline int launch_fail(const TbLaunchCtx *cx, const char *what, const char *detail)
{
    if (cx->err && cx->err_len) snprintf(cx->err, cx->err_len, "%s: %s", what, detail);
    return -2;
}
static inline int check_launch(const TbLaunchCtx *cx, const char *what)
{
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : launch_fail(cx, what, cudaGetErrorString(e));
}
// dynamic shared memory above 48 KB is opt-in per kernel
// (done once per kernel and device: the driver call costs a microsecond or two, which small-batch loops would pay per launch)
static inline int opt_in_smem(const TbLaunchCtx *cx, const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static const void *seen_kernel[256];
    static int seen_dev[256], n_seen = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < n_seen; ++i)
            if (seen_kernel[i] == kernel && seen_dev[i] == dev) return 0;
    }
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return launch_fail(cx, "cudaFuncSetAttribute(shared memory)", cudaGetErrorString(e));
    std::lock_guard<std::mutex> lock(mu);
    if (n_seen < 256) { seen_kernel[n_seen] = kernel; seen_dev[n_seen] = dev; ++n_seen; }
    return 0;
}
// Grid of a kernel that loops over its work items: `waves` x (SMs x resident CTAs per SM).  waves = 1 is a strictly
// persistent grid; the tile kernels use 8 (K1) / 16 (K3) -- tiles and envs differ in cost, and handing the hardware
// scheduler more, shorter CTAs evens the SMs out (K1 -4 %, K3 -2.7 %, random rollout -17 % against waves = 1) at the
// price of a few more table set-ups per SM.  cx->max_ctas (tests only) caps the grid so that every CTA takes several tiles.
static inline unsigned grid_for(const TbLaunchCtx *cx, int64_t work_items, int per_block, int blocks_per_sm, int waves = 1)
{
    const int64_t need = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)cx->sm_count * blocks_per_sm * waves;        // a multiple of the SM count
    if (cx->max_ctas > 0 && cx->max_ctas < cap) cap = cx->max_ctas;
    return (unsigned)(need < 1 ? 1 : (need < cap ? need : cap));
}
// Tile configuration by batch size: the throughput configuration (cfg 0, 256-env tiles, thread per env) once there is
// at least one tile per SM; below that 128-env tiles (`mid`); and for batches of at most cx->small_groups (default 4)
// 32-env groups per SM the small-batch configuration 4 (32 envs and 4 warps per CTA: shortest critical path).
//   K1 cfg  0: 256-thread CTAs, 3 per SM, <= 85 registers (default)   2: 256 x 2, 128 registers   3: 128 x 5
//   K3 cfg  0: 256 x 3, 80 registers (default)   2: 128 x 4   3: 128 x 5   7: 256 x 2, 125 registers
//   both    4: 32 envs x 128 threads   5: 64 envs x 256 threads   (several warps per 32-env group, small batches)
static inline int small_batch_cfg(const TbLaunchCtx *cx, int64_t n_env, int mid)
{
    if ((n_env + 31) / 32 <= (int64_t)cx->small_groups * cx->sm_count) return 4;
    return (n_env + 255) / 256 < cx->sm_count ? mid : 0;
}
static inline F8 f8_from(const float *p, float dflt)
{
    F8 r;
    for (int i = 0; i < 8; ++i) r.v[i] = p ? p[i] : dflt;
    return r;
}

template <int C, int R>
struct ShapeOps {
    using S = Shape<C, R>;
    static StateView view(const void *base, int64_t n_env) { return make_view<C, R>(base, n_env); }

    static int reset(const TbLaunchCtx *cx, void *state, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
                     const uint8_t *tape, const uint8_t *mask)
    {
        k_reset<C, R><<<(unsigned)((n_env + 255) / 256), 256, 0, (cudaStream_t)cx->stream>>>(
            view(state, n_env), env_offset, seed, piece_set, tape, mask);
        return check_launch(cx, "tb_reset");
    }

    static int afterstates(const TbLaunchCtx *cx, const void *state, int64_t n_env, void *feats_out, uint64_t *valid_out,
                           int32_t *count_out, int a_stride, const float *directions, int flags)
    {
        typedef void (*kern_t)(StateView, float *, unsigned long long *, int *, int, F8, int);
        const F8 dirs = f8_from(directions, 1.0f);
        // small batches (fewer 256-env tiles than SMs) use 128-env tiles: twice the CTAs, half the per-tile latency
        const int cfg = cx->k1_cfg >= 0 ? cx->k1_cfg : small_batch_cfg(cx, n_env, 3);
        kern_t kern; size_t smem; int tile, minb, threads = 0;
        const int fmt = (flags & TB_FLAG_FEATS_I16) ? kFmtI16 : (directions ? kFmtF32Dirs : kFmtF32);
#define TB_K1_KERNEL(...) (fmt == kFmtI16 ? k_afterstates<C, R, kFmtI16, __VA_ARGS__> : fmt == kFmtF32Dirs ? \
                           k_afterstates<C, R, kFmtF32Dirs, __VA_ARGS__> : k_afterstates<C, R, kFmtF32, __VA_ARGS__>)
        if (cfg == 4) { tile = 32; threads = 128; minb = 4; smem = sizeof(CtaSmem<C, R, 32>); kern = TB_K1_KERNEL(32, 4, 128); }
        else if (cfg == 5) { tile = 64; threads = 256; minb = 2; smem = sizeof(CtaSmem<C, R, 64>); kern = TB_K1_KERNEL(64, 2, 256); }
        else if (cfg == 3) { tile = 128; minb = 5; smem = sizeof(CtaSmem<C, R, 128>); kern = TB_K1_KERNEL(128, 5); }
        else if (cfg == 2) { tile = 256; minb = 2; smem = sizeof(CtaSmem<C, R, 256>); kern = TB_K1_KERNEL(256, 2); }
        else { tile = 256; minb = 3; smem = sizeof(CtaSmem<C, R, 256>); kern = TB_K1_KERNEL(256, 3); }
#undef TB_K1_KERNEL
        if (opt_in_smem(cx, (const void *)kern, smem)) return -2;
        kern<<<grid_for(cx, n_env, tile, minb, 8), threads ? threads : tile, smem, (cudaStream_t)cx->stream>>>(
            view(state, n_env), (float *)feats_out, (unsigned long long *)valid_out, count_out, a_stride, dirs, flags);
        return check_launch(cx, "tb_afterstates");
    }

    static int afterstates_export(const TbLaunchCtx *cx, const void *state, int64_t n_env, float *feats_out,
                                  uint16_t *rows_out, uint8_t *heights_out, int32_t *info_out, int a_stride)
    {
        const int64_t total = n_env * a_stride;
        k_afterstates_export<C, R><<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)cx->stream>>>(
            view(state, n_env), feats_out, rows_out, heights_out, info_out, a_stride);
        return check_launch(cx, "tb_afterstates_export");
    }

    static int step(const TbLaunchCtx *cx, void *state, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
                    const int32_t *actions, const uint8_t *tape, float *obs, int32_t *reward, uint8_t *done,
                    int32_t *lines, int32_t *status, int flags)
    {
        // k2_cfg (tuning): 0 = 256 envs x 4 CTAs per SM (default), 1 = 128 x 8, 2 = 256 x 3
        const F8 one = f8_from(nullptr, 1.0f);
        cudaStream_t st = (cudaStream_t)cx->stream;
        // small batches (fewer 256-env tiles than SMs): 128-env CTAs, twice as many SMs busy
        if (cx->k2_cfg == 1 || (cx->k2_cfg < 0 && (n_env + 255) / 256 < cx->sm_count))
            k_step<C, R, 128, 8><<<(unsigned)((n_env + 127) / 128), 128, 0, st>>>(
                view(state, n_env), env_offset, seed, piece_set, actions, tape, obs, reward, done, lines, status, flags, one);
        else if (cx->k2_cfg == 2)
            k_step<C, R, 256, 3><<<(unsigned)((n_env + 255) / 256), 256, 0, st>>>(
                view(state, n_env), env_offset, seed, piece_set, actions, tape, obs, reward, done, lines, status, flags, one);
        else if (cx->k2_cfg == 5)
            k_step_tpe<C, R><<<(unsigned)((n_env + 127) / 128), 128, 0, st>>>(
                view(state, n_env), env_offset, seed, piece_set, actions, tape, obs, reward, done, lines, status, flags, one);
        else if (cx->k2_cfg == 3)
            k_step<C, R, 64, 16><<<(unsigned)((n_env + 63) / 64), 64, 0, st>>>(
                view(state, n_env), env_offset, seed, piece_set, actions, tape, obs, reward, done, lines, status, flags, one);
        else if (cx->k2_cfg == 4)
            k_step<C, R, 32, 24><<<(unsigned)((n_env + 31) / 32), 32, 0, st>>>(
                view(state, n_env), env_offset, seed, piece_set, actions, tape, obs, reward, done, lines, status, flags, one);
        else
            k_step<C, R, 256, 4><<<(unsigned)((n_env + 255) / 256), 256, 0, st>>>(
                view(state, n_env), env_offset, seed, piece_set, actions, tape, obs, reward, done, lines, status, flags, one);
        return check_launch(cx, "tb_step");
    }

    static int rollout(const TbLaunchCtx *cx, void *state, int64_t n_env, int64_t env_offset, uint64_t seed, int piece_set,
                       int n_steps, int policy, const float *weights, int64_t *stats, int no_reset, const uint8_t *tape,
                       int tape_stride)
    {
        cudaStream_t st = (cudaStream_t)cx->stream;
        if (policy == TB_POLICY_RANDOM) {
            k_rollout_random<C, R><<<grid_for(cx, n_env, 128, 6, 5), 128, 0, st>>>(
                view(state, n_env), env_offset, seed, piece_set, n_steps, stats, no_reset, tape, tape_stride);
            return check_launch(cx, "tb_rollout");
        }
        typedef void (*kern_t)(StateView, int64_t, uint64_t, int, int, F8, int64_t *, int, const uint8_t *, int);
        kern_t kern; size_t smem; int tile, minb, threads = 0;
        const int cfg = cx->k3_cfg >= 0 ? cx->k3_cfg : small_batch_cfg(cx, n_env, 2);
        if (cfg == 4) { tile = 32; threads = 128; minb = 4; kern = k_rollout_greedy<C, R, 32, 4, 128>;
            smem = ((sizeof(CtaSmem<C, R, 32>) + 15) & ~(size_t)15) + sizeof(BestSmem<32, 128>); }
        else if (cfg == 5) { tile = 64; threads = 256; minb = 2; kern = k_rollout_greedy<C, R, 64, 2, 256>;
            smem = ((sizeof(CtaSmem<C, R, 64>) + 15) & ~(size_t)15) + sizeof(BestSmem<64, 256>); }
        else if (cfg == 3) { tile = 128; minb = 5; kern = k_rollout_greedy<C, R, 128, 5>;
            smem = ((sizeof(CtaSmem<C, R, 128>) + 15) & ~(size_t)15) + sizeof(BestSmem<128>); }
        else if (cfg == 2) { tile = 128; minb = 4; kern = k_rollout_greedy<C, R, 128, 4>;
            smem = ((sizeof(CtaSmem<C, R, 128>) + 15) & ~(size_t)15) + sizeof(BestSmem<128>); }
        else if (cfg == 7) { tile = 256; minb = 2; kern = k_rollout_greedy<C, R, 256, 2>;
            smem = ((sizeof(CtaSmem<C, R, 256>) + 15) & ~(size_t)15) + sizeof(BestSmem<256>); }
        else { tile = 256; minb = 3; kern = k_rollout_greedy<C, R, 256, 3>;   /* 80 registers: the spills sit at the phase
            boundaries (once per env and step), none in phase B; 3 CTAs per SM beat 2 at 125 registers by 5 % */
            smem = ((sizeof(CtaSmem<C, R, 256>) + 15) & ~(size_t)15) + sizeof(BestSmem<256>); }
        if (opt_in_smem(cx, (const void *)kern, smem)) return -2;
        kern<<<grid_for(cx, n_env, tile, minb, 16), threads ? threads : tile, smem, st>>>(
            view(state, n_env), env_offset, seed, piece_set, n_steps, f8_from(weights, 0.0f), stats, no_reset, tape,
            tape_stride);
        return check_launch(cx, "tb_rollout");
    }

    static int fork(const TbLaunchCtx *cx, const void *parent, int64_t n_env, void *child, int a_stride, int n_forks,
                    uint64_t seed2, int64_t child_offset, int piece_set, const uint8_t *tape, int tape_stride)
    {
        const int64_t n_child = n_env * a_stride * n_forks;
        k_fork<C, R><<<(unsigned)((n_child + 127) / 128), 128, 0, (cudaStream_t)cx->stream>>>(
            view(parent, n_env), view(child, n_child), a_stride, n_forks, seed2, child_offset, piece_set, tape, tape_stride);
        return check_launch(cx, "tb_rollout_values(fork)");
    }

    static int export_boards(const TbLaunchCtx *cx, const void *state, int64_t n_env, int64_t first, int64_t count,
                             uint16_t *rows_out, uint8_t *heights_out, uint8_t *piece_out)
    {
        k_export<C, R><<<(unsigned)((count + 127) / 128), 128, 0, (cudaStream_t)cx->stream>>>(
            view(state, n_env), first, count, rows_out, heights_out, piece_out);
        return check_launch(cx, "tb_export_boards");
    }

    static int import_boards(const TbLaunchCtx *cx, void *state, int64_t n_env, int64_t first, int64_t count,
                             const uint16_t *rows_in, const uint8_t *piece_in)
    {
        k_import<C, R><<<(unsigned)((count + 127) / 128), 128, 0, (cudaStream_t)cx->stream>>>(
            view(state, n_env), first, count, rows_in, piece_in);
        return check_launch(cx, "tb_import_boards");
    }

    static int eval_states(const TbLaunchCtx *cx, int64_t n, const uint16_t *rows_in, const int32_t *params,
                           uint16_t *rows_out, uint8_t *heights_out, int32_t *info_out, float *feats_out)
    {
        k_eval_states<C, R><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)cx->stream>>>(
            n, rows_in, params, rows_out, heights_out, info_out, feats_out);
        return check_launch(cx, "tb_eval_states");
    }

    static const TbShapeVT *vt()
    {
        static const TbShapeVT t = {
            TB_SHAPE_ABI, C, R, (size_t)(16 * (S::NB + 1) + 8),
            &reset, &afterstates, &afterstates_export, &step, &rollout, &fork, &export_boards, &import_boards, &eval_states,
        };
        return &t;
    }
};

}  // namespace tb
