// tb_core.cuh -- bit-parallel Tetris afterstate/step math shared by every sm_100a kernel.
//
// Everything here is __host__ __device__ so the exact code the kernels run can also be
// compiled with g++ and checked against the CPU oracle without a GPU (tests/hostcheck).
// It is NOT a CPU product path: tetris_b200 only ever calls the CUDA kernels.
//
// Board model: C columns x R legal rows + 4 buffer rows (game.py:56),
// row 0 = bottom.  Compute form is one 32-bit mask per column (bit r = cell (r, c)); the
// HBM form is one uint16 mask per row, eight rows per 128-bit word (see tb_kernels.cu).
//
// Reference semantics restated in bit-parallel form (file:line into the reference):
//   piece x orientation tables ............ tetromino.py:33-576   (kOri / kPiece below)
//   hard drop by column heights ........... tetromino.py e.g. :349,:365 (anchor_row = max(h - bottom))
//   clear_lines_jitted .................... state.py:121-143      (full = AND of column masks)
//   check_terminal ........................ state.py:111-117      (any cell in row R after clearing)
//   get_feature_values_jitted ............. state.py:175-280      (eval_full / eval_fast)
//   calc_bcts_features .................... state.py:97-107       (landing height, eroded)
//   TetrominoSampler (bag) ................ tetromino.py:12-22    (bag_draw; per-env counter RNG)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TB_HD __host__ __device__ __forceinline__
#else
#define TB_HD inline
#endif

// Bounds / precondition checks of the kernels: compiled in with -DTB_DEBUG_CHECKS (device-side assert -> the launch
// fails with "device-side assert triggered"), compiled out otherwise.  compute-sanitizer is not available on the GPU
// pool, so the parity suite is run once per round against a library built with the checks (profiles/README.md).
#if defined(TB_DEBUG_CHECKS)
#include <assert.h>
#define TB_CHECK(cond) assert(cond)
#else
#define TB_CHECK(cond) ((void)0)
#endif

namespace tb {

// ---------------------------------------------------------------------------------------------
// bit helpers
// ---------------------------------------------------------------------------------------------
TB_HD int popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
TB_HD int clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
TB_HD int ctz32(uint32_t x)  // x != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// Hole depth of one column (state.py:216): for every vertical run of holes, the filled cells above it.  x = the column,
// t = the top cell of every hole run.  The first run is handled without a branch, the others in a loop: most columns
// have at most one, and the purely data-dependent loop ran with 6-9 of 32 lanes (10 % of K2's instructions).  Measured:
// K1 -1.1 %, K2 -1.5 %; two branch-free runs cost K3 2 % (profiles/README.md, r2w).
TB_HD int hole_depth_of(uint32_t x, uint32_t t)
{
    const int r1 = t ? ctz32(t) + 1 : 31;                        // no run at all: x >> 31 = 0 (columns are < 2^31)
    int hd = popc32(x >> r1);
    t &= t - 1u;
    while (t) { hd += popc32(x >> (ctz32(t) + 1)); t &= t - 1u; }
    return hd;
}
TB_HD uint32_t mask_lo(int k) { return (1u << k) - 1u; }          // k in [0, 31]
TB_HD int height_of(uint32_t col) { return 32 - clz32(col); }      // 1 + highest filled row, 0 if empty
TB_HD int imax(int a, int b) { return a > b ? a : b; }
TB_HD int imin(int a, int b) { return a < b ? a : b; }

// cumulative-wells run sum (state.py:227-229,262-272): every maximal vertical run of well
// cells of length L adds 1+2+..+L.  sum_k popc(W_k), W_0 = W, W_{k+1} = W_k & (W_k >> 1).
TB_HD int run_sum(uint32_t w)
{
    int s = 0;
    while (w) { s += popc32(w); w &= w >> 1; }
    return s;
}

// ---------------------------------------------------------------------------------------------
// piece x orientation tables (tetromino.py).  Global piece ids: 0..6 = the 7-piece order of
// game.py:41-47 (Straight, RCorner, LCorner, Square, SnakeR, SnakeL, T); 7,8 = the default set
// of game.py:38-39 (ThreeL, ThreeLine).
// ---------------------------------------------------------------------------------------------
constexpr int kNumPieces = 9;
constexpr int kNumOris = 25;
// piece byte of an env that takes no further part: it has no afterstates and no kernel steps it
constexpr int kPieceDead = 0xFF;      // game over without reset (rollout forks: return -1, game.py:138-145), or a
                                      // terminal board imported from the host
constexpr int kPieceVoid = 0xFE;      // fork of an illegal / non-existent action; an id that names no piece
struct F8 { float v[8]; };            // eight per-feature floats (directions, policy weights) passed by value

// Orientation descriptor, one 32-bit word:
//   [0:3)  w       width in columns
//   [3:23) 4 x 5b  per piece column dx: bottom offset (2b) | cell count (3b) << 2   (count 0 = absent)
//   [23:26) chg    len(changed_lines)
//   [26:28) bonus2 2 * landing_height_bonus
//   [28:31) ph     piece height in rows
struct OriDef { int w; int n; int cell[4][2]; int chg; int bonus2; };

constexpr uint32_t pack_ori(OriDef d)
{
    uint32_t v = (uint32_t)d.w;
    int ph = 0;
    for (int dx = 0; dx < 4; ++dx) {
        int bot = 9, top = -1;
        for (int i = 0; i < d.n; ++i)
            if (d.cell[i][0] == dx) {
                if (d.cell[i][1] < bot) bot = d.cell[i][1];
                if (d.cell[i][1] > top) top = d.cell[i][1];
            }
        if (top >= 0) {
            v |= (uint32_t)(bot | ((top - bot + 1) << 2)) << (3 + 5 * dx);
            if (top + 1 > ph) ph = top + 1;
        }
    }
    return v | (uint32_t)d.chg << 23 | (uint32_t)d.bonus2 << 26 | (uint32_t)ph << 28;
}

// Enumeration order = reference action order: per piece 1-2 column loops; inside a loop the
// orientations are emitted interleaved per column (tetromino.py e.g. :347-378).
#define TB_ORI_TABLE(X)                                                      \
    /* Straight :44-57, :60-74 */                                            \
    X({1, 4, {{0,0},{0,1},{0,2},{0,3}}, 4, 3})                               \
    X({4, 4, {{0,0},{1,0},{2,0},{3,0}}, 1, 0})                               \
    /* RCorner :431-445, :447-460, :465-478, :480-494 */                     \
    X({3, 4, {{0,0},{1,0},{2,0},{2,1}}, 1, 1})                               \
    X({3, 4, {{0,0},{0,1},{1,1},{2,1}}, 2, 1})                               \
    X({2, 4, {{0,2},{1,0},{1,1},{1,2}}, 3, 2})                               \
    X({2, 4, {{0,0},{0,1},{0,2},{1,0}}, 1, 2})                               \
    /* LCorner :511-525, :527-540, :545-559, :561-575 */                     \
    X({3, 4, {{0,0},{1,0},{2,0},{0,1}}, 1, 1})                               \
    X({3, 4, {{2,0},{0,1},{1,1},{2,1}}, 2, 1})                               \
    X({2, 4, {{0,0},{0,1},{0,2},{1,2}}, 3, 2})                               \
    X({2, 4, {{0,0},{1,0},{1,1},{1,2}}, 1, 2})                               \
    /* Square :90-103 */                                                     \
    X({2, 4, {{0,0},{1,0},{0,1},{1,1}}, 2, 1})                               \
    /* SnakeR :120-135, :138-153 */                                          \
    X({3, 4, {{0,0},{1,0},{1,1},{2,1}}, 1, 1})                               \
    X({2, 4, {{0,1},{0,2},{1,0},{1,1}}, 2, 2})                               \
    /* SnakeL :297-312, :315-330 */                                          \
    X({3, 4, {{1,0},{2,0},{0,1},{1,1}}, 1, 1})                               \
    X({2, 4, {{0,0},{0,1},{1,1},{1,2}}, 2, 2})                               \
    /* T :348-362, :364-378, :383-397, :399-413 */                           \
    X({3, 4, {{0,0},{1,0},{2,0},{1,1}}, 1, 1})                               \
    X({3, 4, {{1,0},{0,1},{1,1},{2,1}}, 2, 1})                               \
    X({2, 4, {{0,1},{1,0},{1,1},{1,2}}, 2, 2})                               \
    X({2, 4, {{0,0},{0,1},{0,2},{1,1}}, 2, 2})                               \
    /* ThreeL :216-230, :232-247, :252-266, :267-281 */                      \
    X({2, 3, {{0,0},{1,0},{1,1},{0,0}}, 1, 1})                               \
    X({2, 3, {{0,0},{0,1},{1,1},{0,0}}, 2, 1})                               \
    X({2, 3, {{0,1},{1,0},{1,1},{0,0}}, 2, 1})                               \
    X({2, 3, {{0,0},{0,1},{1,0},{0,0}}, 1, 1})                               \
    /* ThreeLine :168-181, :184-198 */                                       \
    X({1, 3, {{0,0},{0,1},{0,2},{0,0}}, 3, 2})                               \
    X({3, 3, {{0,0},{1,0},{2,0},{0,0}}, 1, 0})

// Piece word: [0:2) n0 orientations in loop 0, [2:5) w0, [5:7) n1, [7:10) w1, [10:16) first ori index.
constexpr uint32_t pack_piece(int n0, int w0, int n1, int w1, int base)
{
    return (uint32_t)n0 | (uint32_t)w0 << 2 | (uint32_t)n1 << 5 | (uint32_t)w1 << 7 | (uint32_t)base << 10;
}
#define TB_PIECE_TABLE(X)                                                    \
    X(pack_piece(1, 1, 1, 4, 0))   /* 0 Straight  */                         \
    X(pack_piece(2, 3, 2, 2, 2))   /* 1 RCorner   */                         \
    X(pack_piece(2, 3, 2, 2, 6))   /* 2 LCorner   */                         \
    X(pack_piece(1, 2, 0, 1, 10))  /* 3 Square    */                         \
    X(pack_piece(1, 3, 1, 2, 11))  /* 4 SnakeR    */                         \
    X(pack_piece(1, 3, 1, 2, 13))  /* 5 SnakeL    */                         \
    X(pack_piece(2, 3, 2, 2, 15))  /* 6 T         */                         \
    X(pack_piece(2, 2, 2, 2, 19))  /* 7 ThreeL    */                         \
    X(pack_piece(1, 1, 1, 3, 23))  /* 8 ThreeLine */

#define TB_X_ORI(...) pack_ori(OriDef __VA_ARGS__),
#define TB_X_PIECE(v) v,
static constexpr uint32_t kOriHost[kNumOris] = { TB_ORI_TABLE(TB_X_ORI) };
static constexpr uint32_t kPieceHost[kNumPieces] = { TB_PIECE_TABLE(TB_X_PIECE) };

// piece sets (game.py:38-39 default, :41-47 seven pieces): local index -> global id
constexpr int kSetSize[2] = { 2, 7 };
TB_HD int set_piece(int piece_set, int idx) { return piece_set == 0 ? 7 + idx : idx; }
TB_HD int set_size(int piece_set) { return piece_set == 0 ? 2 : 7; }

TB_HD int desc_w(uint32_t d) { return (int)(d & 7u); }
TB_HD int desc_bot(uint32_t d, int dx) { return (int)((d >> (3 + 5 * dx)) & 3u); }
TB_HD int desc_len(uint32_t d, int dx) { return (int)((d >> (5 + 5 * dx)) & 7u); }
TB_HD int desc_chg(uint32_t d) { return (int)((d >> 23) & 7u); }
TB_HD int desc_bonus2(uint32_t d) { return (int)((d >> 26) & 3u); }
TB_HD int desc_ph(uint32_t d) { return (int)((d >> 28) & 7u); }

// number of afterstates of a piece on C columns
TB_HD int piece_num_slots(uint32_t pw, int C)
{
    const int n0 = pw & 3, w0 = (pw >> 2) & 7, n1 = (pw >> 5) & 3, w1 = (pw >> 7) & 7;
    return n0 * (C - w0 + 1) + n1 * (C - w1 + 1);
}
// enumeration slot -> (orientation table index, anchor column)
TB_HD void slot_to_placement(uint32_t pw, int C, int slot, int &ori, int &c)
{
    const int n0 = pw & 3, w0 = (pw >> 2) & 7, n1 = (pw >> 5) & 3, base = (pw >> 10) & 63;
    const int s0 = n0 * (C - w0 + 1);
    if (slot < s0) { c = slot >> (n0 - 1); ori = base + (slot & (n0 - 1)); }
    else { const int s = slot - s0; c = s >> (n1 - 1); ori = base + n0 + (s & (n1 - 1)); }
}

// ---------------------------------------------------------------------------------------------
// per-env counter-based RNG + shuffled bag (replaces NumPy's global MT19937 behind
// TetrominoSampler, tetromino.py:12-22; restated on the CPU in oracle/tetris_oracle.c).
// ---------------------------------------------------------------------------------------------
TB_HD uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
TB_HD uint64_t env_key(uint64_t seed, uint64_t env) { return mix64(seed + 0x9E3779B97F4A7C15ULL * (env + 1)); }
TB_HD uint32_t rng32(uint64_t key, uint32_t ctr, uint32_t stream)
{
    return (uint32_t)(mix64(key ^ (((uint64_t)stream << 32) | (uint64_t)ctr)) >> 32);
}
TB_HD uint32_t bounded(uint32_t r, uint32_t k) { return (uint32_t)(((uint64_t)r * (uint64_t)k) >> 32); }

// Draw the next local piece index: refill the bag when empty, pick the j-th remaining piece.
TB_HD int bag_draw(int n_set, uint64_t key, uint32_t &bag, uint32_t &draws)
{
    if (bag == 0) bag = (1u << n_set) - 1u;
    const uint32_t j = bounded(rng32(key, draws, 0u), (uint32_t)popc32(bag));
    draws += 1;
    uint32_t b = bag;
    for (uint32_t i = 0; i < j; ++i) b &= b - 1;
    TB_CHECK(b != 0u);
    const int idx = ctz32(b);
    bag &= ~(1u << idx);
    return idx;
}

// ---------------------------------------------------------------------------------------------
// row-mask <-> column-mask transposition.
// HBM holds rows packed two per 32-bit word (row 2k in the low half), 8 rows per 128-bit plane.
// ---------------------------------------------------------------------------------------------
TB_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
#endif
}
// 8x8 bit-matrix transpose of the 64-bit matrix (lo = rows 0-3, hi = rows 4-7, one byte per row).
TB_HD void transpose8(uint32_t &lo, uint32_t &hi)
{
    uint32_t t;
    t = (lo ^ (lo >> 7)) & 0x00AA00AAu; lo ^= t ^ (t << 7);
    t = (hi ^ (hi >> 7)) & 0x00AA00AAu; hi ^= t ^ (t << 7);
    t = (lo ^ (lo >> 14)) & 0x0000CCCCu; lo ^= t ^ (t << 14);
    t = (hi ^ (hi >> 14)) & 0x0000CCCCu; hi ^= t ^ (t << 14);
    t = (hi & 0xF0F0F0F0u) | ((lo >> 4) & 0x0F0F0F0Fu);
    lo = (lo & 0x0F0F0F0Fu) | ((hi << 4) & 0xF0F0F0F0u);
    hi = t;
}

template <int C, int R>
struct Shape {
    static constexpr int N = R + 4;                      // stored rows (game.py:56)
    static constexpr int NB = (N + 7) / 8;               // 8-row blocks = 128-bit planes per env
    static constexpr int NW = NB * 4;                    // packed 32-bit words per env
    static constexpr uint32_t ALL = (N >= 32) ? 0xFFFFFFFFu : ((1u << N) - 1u);
    static constexpr uint32_t FULLROW = (1u << C) - 1u;
    // uint16 row masks hold 16 columns; 32-bit column masks hold R + 4 stored rows.  R + 4 <= 31 (not 32) keeps every
    // shift amount in the code below 32 -- a terminal afterstate's column can be R + 4 cells tall.
    static_assert(C >= 4 && C <= 16, "supported widths: 4..16 columns");
    static_assert(N <= 31 && R >= 4, "supported heights: 4..27 rows (+4 buffer rows)");
};

// rows (packed words w[NW]) -> column masks col[C]
template <int C, int R>
TB_HD void rows_to_cols(const uint32_t *w, uint32_t *col)
{
    using S = Shape<C, R>;
    uint32_t lo[2][S::NB], hi[2][S::NB];
#pragma unroll
    for (int b = 0; b < S::NB; ++b) {
        // block b = rows 8b..8b+7 = words 4b..4b+3; half 0 = columns 0-7, half 1 = columns 8-15
        lo[0][b] = prmt(w[4 * b], w[4 * b + 1], 0x6420);
        hi[0][b] = prmt(w[4 * b + 2], w[4 * b + 3], 0x6420);
        transpose8(lo[0][b], hi[0][b]);
        if (C > 8) {
            lo[1][b] = prmt(w[4 * b], w[4 * b + 1], 0x7531);
            hi[1][b] = prmt(w[4 * b + 2], w[4 * b + 3], 0x7531);
            transpose8(lo[1][b], hi[1][b]);
        }
    }
    // after the transpose byte j of (lo,hi) holds column j's bits for the block's 8 rows
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int half = c >> 3, j = c & 7;
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < S::NB; ++b) {
            const uint32_t src = (j < 4) ? lo[half][b] : hi[half][b];
            v |= ((src >> (8 * (j & 3))) & 0xffu) << (8 * b);
        }
        col[c] = v;
    }
}

// column masks -> rows (packed words)
template <int C, int R>
TB_HD void cols_to_rows(const uint32_t *col, uint32_t *w)
{
    using S = Shape<C, R>;
#pragma unroll
    for (int b = 0; b < S::NB; ++b) {
        uint32_t lo[2] = {0, 0}, hi[2] = {0, 0};
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int half = c >> 3, j = c & 7;
            const uint32_t byte = (col[c] >> (8 * b)) & 0xffu;
            if (j < 4) lo[half] |= byte << (8 * j); else hi[half] |= byte << (8 * (j - 4));
        }
        transpose8(lo[0], hi[0]);          // now byte i = row (8b+i) bits for columns 0-7
        if (C > 8) transpose8(lo[1], hi[1]);
        // word 4b+k holds rows 8b+2k (low half) and 8b+2k+1 (high half)
        w[4 * b + 0] = prmt(lo[0], lo[1], 0x5140);
        w[4 * b + 1] = prmt(lo[0], lo[1], 0x7362);
        w[4 * b + 2] = prmt(hi[0], hi[1], 0x5140);
        w[4 * b + 3] = prmt(hi[0], hi[1], 0x7362);
    }
}

// ---------------------------------------------------------------------------------------------
// from-scratch evaluation of the six board features (state.py:175-280), general form: valid for
// any board with consistent heights, terminal or not (walls: all-ones mask, height R).
// out6 = [rows_with_holes, column_transitions, holes, cumulative_wells, row_transitions, hole_depth]
// ---------------------------------------------------------------------------------------------
template <int C, int R>
TB_HD void eval_full(const uint32_t *col, int *out6)
{
    using S = Shape<C, R>;
    int holes = 0, ct = 0, hd = 0, wells = 0, rt = 0;
    uint32_t hm = 0;
    uint32_t L = S::ALL;
    int hL = R;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const uint32_t x = col[c];
        const int h = height_of(x);
        const uint32_t Rt = (c + 1 < C) ? col[c + 1] : S::ALL;
        const int hR = (c + 1 < C) ? height_of(col[c + 1]) : R;
        const uint32_t mh = mask_lo(h);
        const uint32_t hole = ~x & mh;
        holes += popc32(hole);                                   // state.py:213
        hm |= hole;                                              // :215
        uint32_t t = hole & (x >> 1);                            // top cell of every vertical hole run
        ct += 1 + 2 * popc32(t);                                 // :194,:219-220,:242-243
        while (t) {                                              // :216 filled cells above the run
            const int r = ctz32(t);
            hd += popc32(x >> (r + 1));
            t &= t - 1;
        }
        const int lim = imax(h, imin(hL, hR));                   // :258-260
        wells += run_sum(L & Rt & ~x & mask_lo(lim));            // :222-233,:262-272
        if (h > 0) rt += imax(0, hL - h) + popc32((x ^ L) & mh); // :203-204,:225-226,:246-248
        else rt += popc32(L & mask_lo(hL));                      // :254
        L = x; hL = h;
    }
    rt += R - popc32(col[C - 1]);                                // :190 right wall
    out6[0] = popc32(hm); out6[1] = ct; out6[2] = holes; out6[3] = wells; out6[4] = rt; out6[5] = hd;
}

// remove the rows in `full` from every column (rows above shift down): state.py:126-131
template <int C>
TB_HD void clear_rows(uint32_t *col, uint32_t full)
{
    while (full) {
        const int r = 31 - clz32(full);                          // highest first: lower indices stay valid
        const uint32_t lowm = mask_lo(r);
#pragma unroll
        for (int c = 0; c < C; ++c) col[c] = (col[c] & lowm) | ((col[c] >> 1) & ~lowm);
        full &= ~(1u << r);
    }
}

// Result of evaluating one placement.
struct Eval {
    float f[8];          // [rows_with_holes, column_transitions, holes, landing_height, cumulative_wells,
                         //  row_transitions, eroded, hole_depth]  (game.py:10-18)
    int a;               // anchor_row (pre-clear)
    uint32_t full;       // cleared rows (absolute row mask)
    int terminal;        // State.terminal_state
};

// hard-drop row of orientation `d` anchored at column c (tetromino.py: anchor_row = max(h - bottom)).
// Static indexing over the board's columns so `col` can stay in registers.
// Up to 12 columns the loops over the board's columns are branch-free: what the piece contributes to column k is cut out
// of a 64-bit word that holds the piece's four per-column fields shifted to the anchor column (5-bit fields for the
// bottom offsets, 4-bit fields for the cells).  The per-column `if (k - c < 4)` form ran each column's body with the 2-3
// lanes of a warp whose piece covers that column (a quarter of the random rollout's instructions at 9-15 of 32 lanes).
template <int C>
TB_HD int anchor_from_cols(const uint32_t *col, uint32_t d, int c)
{
    if (C <= 12) {
        uint32_t nb5 = 0u;                                       // per piece column: 31 - bot (a column without cells: 0)
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
            const uint32_t f = (d >> (3 + 5 * dx)) & 31u;
            nb5 |= ((f >> 2) ? 31u - (f & 3u) : 0u) << (5 * dx);
        }
        const unsigned long long q = (unsigned long long)nb5 << (5 * c);   // fields beyond column C-1 are empty
        int a = 0;
#pragma unroll
        for (int k = 0; k < C; ++k)
            a = imax(a, height_of(col[k]) + (int)((uint32_t)(q >> (5 * k)) & 31u) - 31);   // h - bot, or below zero
        return a;
    }
    int a = 0;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const unsigned dx = (unsigned)(k - c);
        if (dx < 4u) {
            const uint32_t f = (d >> (3 + 5 * dx)) & 31u;
            if (f >> 2) a = imax(a, height_of(col[k]) - (int)(f & 3u));
        }
    }
    return a;
}

// The cells of orientation `d`, per piece column relative to the anchor row, as four 4-bit fields
TB_HD uint32_t piece_cells4(uint32_t d)
{
    uint32_t p4 = 0u;
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) {
        const uint32_t f = (d >> (3 + 5 * dx)) & 31u;
        p4 |= (mask_lo((int)(f >> 2)) << (f & 3u)) << (4 * dx);
    }
    return p4;
}

// Place orientation `d` at column c on `col` (in place), clear lines, report a/full/terminal.
// Returns the number of piece cells in the cleared rows (for the eroded feature).
template <int C, int R>
TB_HD int place_and_clear(uint32_t *col, uint32_t d, int c, int &a_out, uint32_t &full_out, int &terminal)
{
    using S = Shape<C, R>;
    const int a = anchor_from_cols<C>(col, d, c);
    uint32_t full = S::ALL;
    int cells = 0;
    if (C <= 12) {
        const unsigned long long q = (unsigned long long)piece_cells4(d) << (4 * c);
#pragma unroll
        for (int k = 0; k < C; ++k) {
            col[k] |= ((uint32_t)(q >> (4 * k)) & 15u) << a;
            full &= col[k];
        }
        full &= mask_lo(desc_chg(d)) << a;                       // only changed_lines are tested (state.py:122)
        if (full) {
#pragma unroll
            for (int k = 0; k < C; ++k) cells += popc32((((uint32_t)(q >> (4 * k)) & 15u) << a) & full);
            clear_rows<C>(col, full);
        }
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const unsigned dx = (unsigned)(k - c);
            if (dx < 4u) {
                const uint32_t f = (d >> (3 + 5 * dx)) & 31u;
                col[k] |= mask_lo((int)(f >> 2)) << (a + (int)(f & 3u));
            }
            full &= col[k];
        }
        full &= mask_lo(desc_chg(d)) << a;                       // only changed_lines are tested (state.py:122)
        if (full) {
#pragma unroll
            for (int k = 0; k < C; ++k) {
                const unsigned dx = (unsigned)(k - c);
                if (dx < 4u) {
                    const uint32_t f = (d >> (3 + 5 * dx)) & 31u;
                    cells += popc32((mask_lo((int)(f >> 2)) << (a + (int)(f & 3u))) & full);
                }
            }
            clear_rows<C>(col, full);
        }
    }
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < C; ++k) any |= col[k];
    terminal = (int)((any >> R) & 1u);                            // state.py:111-117
    a_out = a; full_out = full;
    return cells;
}

// Slow (general) path: build the afterstate board, clear, evaluate from scratch.
// `col` is the current board (not modified unless out_col == col); out_col (nullable) gets the afterstate.
template <int C, int R, bool PERM = true>
TB_HD void eval_full_tab(const uint32_t *col, const uint32_t *runtab, int *out6);   // below, after the run-sum table

// runtab (nullable): the run-sum table; with it the afterstate must be NON-TERMINAL (eval_full_tab).
// PERM: the table carries its byte-permuted second copy (RunTab::COPIES; K2 stages the plain table only).
template <int C, int R, bool PERM = true>
TB_HD void eval_slow(const uint32_t *col, uint32_t d, int c, Eval &e, uint32_t *out_col, const uint32_t *runtab = nullptr)
{
    uint32_t nc[C];
#pragma unroll
    for (int k = 0; k < C; ++k) nc[k] = col[k];
    const int cells = place_and_clear<C, R>(nc, d, c, e.a, e.full, e.terminal);
    int six[6];
    if (runtab) eval_full_tab<C, R, PERM>(nc, runtab, six);
    else eval_full<C, R>(nc, six);
    const int ncl = popc32(e.full);
    e.f[0] = (float)six[0]; e.f[1] = (float)six[1]; e.f[2] = (float)six[2];
    e.f[3] = (float)(2 * (e.a + 1) + desc_bonus2(d)) * 0.5f;      // state.py:102 (pre-clear anchor)
    e.f[4] = (float)six[3]; e.f[5] = (float)six[4];
    e.f[6] = (float)(cells * ncl);                               // state.py:99-101
    e.f[7] = (float)six[5];
    if (out_col) {
#pragma unroll
        for (int k = 0; k < C; ++k) out_col[k] = nc[k];
    }
}

// ---------------------------------------------------------------------------------------------
// run-sum by table.  A well mask of a non-terminal board has bits only below row R.  It is cut into NCH chunks of
// HB = ceil(R / NCH) bits (1 chunk up to R = 10, 2 up to R = 20, 3 above); the table has 2^HB 32-bit entries, entry
// for a chunk m = run_sum(m) | lead << 8 | trail << 16 | 1 << 24, where trail / lead = length of the run of ones
// touching the chunk's bit 0 / bit HB-1.  A run spanning two chunks adds the cross term of
// (p+q)(p+q+1)/2 = p(p+1)/2 + q(q+1)/2 + p*q, so for two chunks
//     run_sum = rs0 * 1 + lead0 * trail1 + trail0 * 0 + 1 * rs1
// which is ONE dot-product instruction (IDP.4A) on the low entry and a byte-permuted high entry.
// No POPC, no data-dependent loop, no warp vote: the same code on every board.
// ---------------------------------------------------------------------------------------------
template <int R>
struct RunTab {
    static constexpr int NCH = R <= 10 ? 1 : (R <= 20 ? 2 : 3);
    static constexpr int HB = (R + NCH - 1) / NCH;
    static constexpr int SIZE = 1 << HB;
    // two-chunk boards keep a second copy with every entry already byte-permuted into the form the dot product wants
    // of the HIGH chunk ([1, trail, 0, run_sum]): one PRMT less per lookup pair
    static constexpr int COPIES = NCH == 2 ? 2 : 1;
    static constexpr int WORDS = COPIES * SIZE;
    static_assert(HB <= 10 && R <= 27, "run table: chunks of at most 10 rows, boards of at most 27 rows");
};
template <int R>
TB_HD uint32_t run_tab_entry(uint32_t m)
{
    constexpr int HB = RunTab<R>::HB;
    const int rs = run_sum(m);
    const uint32_t z = ~m & mask_lo(HB);                           // the zero bits of the chunk
    const int trail = z ? ctz32(z) : HB;                           // ones below the lowest zero
    const int lead = z ? clz32(z) - (32 - HB) : HB;                // ones above the highest zero
    return (uint32_t)rs | ((uint32_t)lead << 8) | ((uint32_t)trail << 16) | (1u << 24);
}
TB_HD uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
    return c;
#endif
}
// acc + run_sum(w), w < 2^R.
// On the device the table is ALWAYS in shared memory (every kernel stages it) and ALIGNED TO ITS SIZE, so that the
// address of an entry is base | byte offset -- one LOP3 together with the mask of the chunk -- and the load is spelled
// ld.shared on a 32-bit shared-space address.  (With a plain pointer ptxas sometimes keeps the table's base in a vector
// register and adds it per lookup: +1 instruction on each of the ~9 lookups per afterstate.)
template <int R, bool PERM = true>
TB_HD uint32_t run_sum_acc(const uint32_t *tab, uint32_t w, uint32_t acc)
{
    constexpr int HB = RunTab<R>::HB, NCH = RunTab<R>::NCH;
    constexpr uint32_t M = (uint32_t)(RunTab<R>::SIZE - 1);
    TB_CHECK((NCH * HB >= 32) || (w >> (NCH * HB)) == 0u);         // a well mask of a non-terminal board: below row R
#if defined(__CUDA_ARCH__)
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    TB_CHECK((base & (4u * M + 3u)) == 0u);
    auto entry = [&](uint32_t byte_off) {                          // byte_off = 4 * index, already masked
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base | byte_off));
        return v;
    };
    const uint32_t e0 = entry((w << 2) & (4u * M));
    if (NCH == 1) return acc + (e0 & 255u);
    const uint32_t off1 = HB >= 2 ? (w >> (HB - 2)) & (4u * M) : ((w >> HB) & M) << 2;
    if (NCH == 2 && PERM) {                                        // the permuted copy sits SIZE entries further
        uint32_t e1p;
        asm("ld.shared.u32 %0, [%1+%2];" : "=r"(e1p) : "r"(base | off1), "n"(4 * RunTab<R>::SIZE));
        return dp4a_u(e0, e1p, acc);
    }
    const uint32_t e1 = entry(off1);
#else
    const uint32_t e0 = tab[w & M];
    if (NCH == 1) return acc + (e0 & 255u);
    if (NCH == 2 && PERM) return dp4a_u(e0, tab[RunTab<R>::SIZE + ((w >> HB) & M)], acc);
    const uint32_t e1 = tab[(w >> HB) & M];
#endif
    // e1 permuted to bytes [1, trail1, 0, rs1] against e0 = [rs0, lead0, trail0, 1]
    acc = dp4a_u(e0, prmt(e1, 0u, 0x0423), acc);
    if (NCH == 3) {
#if defined(__CUDA_ARCH__)
        const uint32_t e2 = entry((w >> (2 * HB - 2)) & (4u * M));
#else
        const uint32_t e2 = tab[w >> (2 * HB)];
#endif
        const uint32_t lead0 = (e0 >> 8) & 255u, lead1 = (e1 >> 8) & 255u, trail1 = (e1 >> 16) & 255u;
        const uint32_t carry = trail1 == (uint32_t)HB ? lead0 + (uint32_t)HB : lead1;   // run reaching chunk 2 from below
        acc += (e2 & 255u) + carry * ((e2 >> 16) & 255u);
    }
    return acc;
}
template <int R>
TB_HD int run_sum_tab(const uint32_t *tab, uint32_t w) { return (int)run_sum_acc<R>(tab, w, 0u); }

// eval_full for NON-TERMINAL boards (every column height <= R: well masks stay below row R), cumulative wells by table:
// no data-dependent POPC loop, so the lanes of a warp stay together.  Same out6 as eval_full.
template <int C, int R, bool PERM>
TB_HD void eval_full_tab(const uint32_t *col, const uint32_t *runtab, int *out6)
{
    using S = Shape<C, R>;
    int holes = 0, ct = 0, hd = 0, rt = 0;
    uint32_t wells = 0u, hm = 0u, L = S::ALL;
    int hL = R;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const uint32_t x = col[c];
        const int h = height_of(x);
        const uint32_t Rt = (c + 1 < C) ? col[c + 1] : S::ALL;
        const uint32_t mh = mask_lo(h);
        const uint32_t hole = ~x & mh;
        holes += popc32(hole);                                   // state.py:213
        hm |= hole;                                              // :215
        uint32_t t = hole & (x >> 1);                            // top cell of every vertical hole run
        ct += 1 + 2 * popc32(t);                                 // :194,:219-220,:242-243
        hd += hole_depth_of(x, t);                               // :216 filled cells above the run
        wells = run_sum_acc<R, PERM>(runtab, L & Rt & ~x, wells);   // :222-233,:262-272 (heights <= R: no limit mask needed)
        if (h > 0) rt += imax(0, hL - h) + popc32((x ^ L) & mh); // :203-204,:225-226,:246-248
        else rt += popc32(L & mask_lo(hL));                      // :254
        L = x; hL = h;
    }
    rt += R - popc32(col[C - 1]);                                // :190 right wall
    out6[0] = popc32(hm); out6[1] = ct; out6[2] = holes; out6[3] = (int)wells; out6[4] = rt; out6[5] = hd;
}

// Small non-negative int -> float without the conversion unit: the bit pattern 0x4B000000 + v is the float
// 2^23 + v.  The env record keeps its totals already biased, so adding a placement's integer delta yields the
// float bits directly and only the subtraction of 2^23 remains.
constexpr uint32_t kFloatBias = 0x4B000000u;
TB_HD float unbias(uint32_t bits)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(bits) - 8388608.0f;
#else
    float f;
    __builtin_memcpy(&f, &bits, 4);
    return f - 8388608.0f;
#endif
}
TB_HD float u2f(int v) { return unbias(kFloatBias + (uint32_t)v); }

// ---------------------------------------------------------------------------------------------
// Per-env record for incremental evaluation.  Built once per env (build_env), read by every placement of that
// env (eval_placement).  Flat uint32 words so it can live in shared memory; neighbour arrays are padded with
// wall / zero sentinels so a placement at any column needs no range checks.
// ---------------------------------------------------------------------------------------------
template <int C, int R>
struct Env {
    static constexpr int COLX = 0;                         // C+4 words: [k] = column k-2; walls (ALL) at -1 and C, 0 at -2, C+1
    static constexpr int FQ = COLX + C + 4;                // 4 words: [q-1] = rows with exactly q empty cells, q = 1..4
    static constexpr int H8 = FQ + 4;                      // C+2 bytes: [k] = height of column k-1 (walls: R)
    static constexpr int NR8 = H8 + (C + 2 + 3) / 4;       // C bytes: hole runs per column; byte C: tallest column
    static constexpr int PW16 = NR8 + (C + 1 + 3) / 4;     // C+3 u16: [i] = wells of columns < clamp(i-1, 0, C)
    static constexpr int PRT16 = PW16 + (C + 3 + 1) / 2;   // C+2 u16: [i] = row transitions of columns < min(i, C)
    static constexpr int TOT = PRT16 + (C + 2 + 1) / 2;    // ct, hd, wells, rt, holes (each + kFloatBias), hole-row mask
    static constexpr int T_CT = TOT + 0, T_HD = TOT + 1, T_WELLS = TOT + 2, T_RT = TOT + 3,
                         T_HOLES = TOT + 4, T_HM = TOT + 5;
    static constexpr int B_HMAX = 4 * NR8 + C;             // byte index of the tallest column's height
    static constexpr int WORDS_RAW = TOT + 6;
    // Record stride in words.  Odd, so that thread t working on record t is conflict-free whatever the field.  For
    // C = 10 the natural size is 43 = 32 + 11: the records of three envs at consecutive positions start 11 banks apart
    // and those of four envs at positions 8 apart start 8 banks apart -- what makes the windows of the tile kernels
    // (3 envs x 9-10 anchor columns, 4 envs x 7-8) free of shared-memory bank conflicts (tb_kernels.cuh, window_slot).
    static constexpr int WORDS = WORDS_RAW | 1;
};

// Builds the record from the columns already stored in rec[COLX + 2 .. COLX + 2 + C).  A rolled loop on purpose: an
// unrolled builder is 8x the code, and instruction-cache footprint matters more than a few extra shared-memory
// accesses (K1 / K3 are instruction-latency bound).  Precondition: every column height is <= R (non-terminal state).
template <int C, int R>
TB_HD void build_env(const uint32_t *runtab, uint32_t *rec)
{
    using S = Shape<C, R>;
    using K = Env<C, R>;
    uint8_t *rb = reinterpret_cast<uint8_t *>(rec);
    uint16_t *rh = reinterpret_cast<uint16_t *>(rec);
    rec[K::COLX + 0] = 0u; rec[K::COLX + 1] = S::ALL;
    rec[K::COLX + C + 2] = S::ALL; rec[K::COLX + C + 3] = 0u;
    rb[4 * K::H8 + 0] = (uint8_t)R; rb[4 * K::H8 + C + 1] = (uint8_t)R;
    rh[2 * K::PW16 + 0] = 0; rh[2 * K::PW16 + 1] = 0; rh[2 * K::PRT16 + 0] = 0;
    int holes = 0, ct = 0, hd = 0, rt = 0, hL = R;
    uint32_t wells = 0;
    uint32_t hm = 0, L = S::ALL, any = 0;
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;                       // bit-sliced count of empty cells per row (c3: >= 8)
    uint32_t x = rec[K::COLX + 2];
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        const uint32_t Rt = rec[K::COLX + 3 + c];                  // column c + 1, or the right wall sentinel
        {
            const uint32_t e = ~x;
            const uint32_t k0 = c0 & e; c0 ^= e;
            const uint32_t k1 = c1 & k0; c1 ^= k0;
            const uint32_t k2 = c2 & k1; c2 ^= k1;
            c3 |= k2;
        }
        const int h = height_of(x);
        const uint32_t mh = mask_lo(h);
        const uint32_t hole = ~x & mh;
        holes += popc32(hole);
        hm |= hole;
        any |= x;
        uint32_t t = hole & (x >> 1);
        const int nr = popc32(t);
        ct += 1 + 2 * nr;
        hd += hole_depth_of(x, t);
        wells = run_sum_acc<R>(runtab, L & Rt & ~x, wells);
        if (h > 0) rt += imax(0, hL - h) + popc32((x ^ L) & mh);
        else rt += popc32(L & mask_lo(hL));
        rb[4 * K::H8 + c + 1] = (uint8_t)h;
        rb[4 * K::NR8 + c] = (uint8_t)nr;
        rh[2 * K::PW16 + c + 2] = (uint16_t)wells;
        rh[2 * K::PRT16 + c + 1] = (uint16_t)rt;
        L = x; hL = h; x = Rt;
    }
    rh[2 * K::PW16 + C + 2] = (uint16_t)wells;
    rh[2 * K::PRT16 + C + 1] = (uint16_t)rt;
    rt += R - popc32(L);                                           // L = column C-1 after the loop
    // Full-row detection without looking at the other columns: every cell of a dropped piece lands on an EMPTY cell
    // (above its column's height), so a changed row becomes full exactly when its number of empty cells equals the
    // number of piece cells the orientation puts into it (pieces_per_changed_row, 1..4).  state.py:122-124 restated.
    const uint32_t lo3 = ~c3;
    rec[K::FQ + 0] = c0 & ~c1 & ~c2 & lo3;
    rec[K::FQ + 1] = ~c0 & c1 & ~c2 & lo3;
    rec[K::FQ + 2] = c0 & c1 & ~c2 & lo3;
    rec[K::FQ + 3] = ~c0 & ~c1 & c2 & lo3;
    rec[K::T_CT] = kFloatBias + (uint32_t)ct; rec[K::T_HD] = kFloatBias + (uint32_t)hd;
    rec[K::T_WELLS] = kFloatBias + wells; rec[K::T_RT] = kFloatBias + (uint32_t)rt;
    rec[K::T_HOLES] = kFloatBias + (uint32_t)holes; rec[K::T_HM] = hm;
    rb[K::B_HMAX] = (uint8_t)height_of(any);
}

// eval_placement status
constexpr int kFastDone = 0;      // features written, afterstate is legal
constexpr int kFastClears = 1;    // a line clears: features need the general path; e.a / e.full / e.terminal are set
constexpr int kFastTerminal = 2;  // no line clears and the piece reaches row R: terminal afterstate

// What a placement of width W at column c reads from the env record, independent of the orientation: shared by
// the orientations of one column loop.
template <int C, int R, int W>
struct Neigh {
    uint32_t y[W + 4];       // columns c-2 .. c+W+1
    uint32_t mh[W + 1];      // mask_lo(height) of columns c .. c+W
    int h[W + 2];            // heights of columns c-1 .. c+W
    int nr[W];               // hole runs of columns c .. c+W-1
    uint32_t wells0, rt0;    // board totals (+ kFloatBias) minus the contributions of the columns the placement can change
};
template <int C, int R, int W>
TB_HD void load_neigh(const uint32_t *rec, int c, Neigh<C, R, W> &nb)
{
    using K = Env<C, R>;
    const uint8_t *rb = reinterpret_cast<const uint8_t *>(rec);
    const uint16_t *rh = reinterpret_cast<const uint16_t *>(rec);
#pragma unroll
    for (int k = 0; k < W + 4; ++k) nb.y[k] = rec[K::COLX + c + k];
#pragma unroll
    for (int k = 0; k < W + 2; ++k) nb.h[k] = (int)rb[4 * K::H8 + c + k];
#pragma unroll
    for (int k = 0; k < W; ++k) nb.nr[k] = (int)rb[4 * K::NR8 + c + k];
#pragma unroll
    for (int k = 0; k < W + 1; ++k) nb.mh[k] = mask_lo(nb.h[1 + k]);
    nb.wells0 = rec[K::T_WELLS] - ((uint32_t)rh[2 * K::PW16 + c + W + 2] - (uint32_t)rh[2 * K::PW16 + c]);
    nb.rt0 = rec[K::T_RT] - ((uint32_t)rh[2 * K::PRT16 + c + W + 1] - (uint32_t)rh[2 * K::PRT16 + c]);
}

// Orientation descriptor decoded into the values eval_neigh uses.  In the kernels the descriptor is warp-uniform
// and is decoded in converged code, so these live in uniform registers and cost no vector-ALU work.
struct OriU {
    int bot[4], len[4], top[4];   // per piece column: lowest cell offset, cell count, bot + len
    uint32_t seg[4];              // mask_lo(len) << bot: the column's cells relative to the anchor row
    uint32_t p2bot[4], p2top[4];  // 1 << bot, 1 << top: mask_lo(a + bot) = (1 << a) * p2bot - 1 is one multiply-add
    uint32_t chgm;                // mask_lo(len(changed_lines))
    int ph;                       // piece height
    uint32_t lh2;                 // 2 + 2 * landing_height_bonus + kFloatBias
    // pieces_per_changed_row takes at most two distinct values qa, qb over an orientation's changed rows:
    // qa1 / qb1 = value - 1 (index of the record's FQ word), ma / mb = the changed rows (relative to the anchor) with
    // that many piece cells.  One value only: qb1 = qa1, mb = 0.
    uint32_t qa1, ma, qb1, mb;
};
constexpr int kOriWords = 32;     // OriU as flat words: bot[4] len[4] top[4] seg[4] p2bot[4] p2top[4] chgm ph lh2 qa1 ma qb1 mb pad
TB_HD OriU decode_ori(uint32_t d)
{
    OriU u;
#pragma unroll
    for (int dx = 0; dx < 4; ++dx) {
        u.bot[dx] = desc_bot(d, dx);
        u.len[dx] = desc_len(d, dx);
        u.top[dx] = u.bot[dx] + u.len[dx];
        u.seg[dx] = mask_lo(u.len[dx]) << u.bot[dx];
        u.p2bot[dx] = 1u << u.bot[dx];
        u.p2top[dx] = 1u << u.top[dx];
    }
    u.chgm = mask_lo(desc_chg(d));
    u.ph = desc_ph(d);
    u.lh2 = 2u + (uint32_t)desc_bonus2(d) + kFloatBias;
    // pieces_per_changed_row of row k = number of piece columns whose cells cover relative row k
    u.qa1 = 0u; u.ma = 0u; u.qb1 = 0u; u.mb = 0u;
    for (int k = 0; k < desc_chg(d); ++k) {
        uint32_t q = 0;
        for (int dx = 0; dx < 4; ++dx) q += (u.seg[dx] >> k) & 1u;
        if (u.ma == 0u || u.qa1 == q - 1u) { u.qa1 = q - 1u; u.ma |= 1u << k; }
        else { u.qb1 = q - 1u; u.mb |= 1u << k; }
    }
    if (u.mb == 0u) u.qb1 = u.qa1;
    return u;
}
static_assert(sizeof(OriU) == 31 * 4, "OriU is 31 words");

// ---------------------------------------------------------------------------------------------
// Compile-time images of the two tables the tile kernels keep in shared memory:
//   OdescImage  : decode_ori() of every orientation, laid out as OriU (kOriWords words each)
//   RunImage<R> : run_tab_entry<R>() of every half mask
// Plain constexpr restatements (no intrinsics), so the kernels can copy instead of compute; tests/hostcheck
// compares them word for word with decode_ori / run_tab_entry.
// ---------------------------------------------------------------------------------------------
struct OdescImage { uint32_t w[kNumOris][kOriWords]; };
constexpr uint32_t cx_mask(int k) { return (1u << k) - 1u; }
constexpr OdescImage make_odesc_image()
{
    OdescImage t{};
    for (int i = 0; i < kNumOris; ++i) {
        const uint32_t d = kOriHost[i];
        for (int dx = 0; dx < 4; ++dx) {
            const uint32_t bot = (d >> (3 + 5 * dx)) & 3u, len = (d >> (5 + 5 * dx)) & 7u, top = bot + len;
            t.w[i][0 + dx] = bot; t.w[i][4 + dx] = len; t.w[i][8 + dx] = top;
            t.w[i][12 + dx] = cx_mask((int)len) << bot; t.w[i][16 + dx] = 1u << bot; t.w[i][20 + dx] = 1u << top;
        }
        t.w[i][24] = cx_mask((int)((d >> 23) & 7u));                 // chgm
        t.w[i][25] = (d >> 28) & 7u;                                 // ph
        t.w[i][26] = 2u + ((d >> 26) & 3u) + kFloatBias;             // lh2
        uint32_t qa1 = 0, ma = 0, qb1 = 0, mb = 0;
        for (uint32_t k = 0; k < ((d >> 23) & 7u); ++k) {
            uint32_t q = 0;
            for (int dx = 0; dx < 4; ++dx) q += (t.w[i][12 + dx] >> k) & 1u;
            if (ma == 0u || qa1 == q - 1u) { qa1 = q - 1u; ma |= 1u << k; }
            else { qb1 = q - 1u; mb |= 1u << k; }
        }
        if (mb == 0u) qb1 = qa1;
        t.w[i][27] = qa1; t.w[i][28] = ma; t.w[i][29] = qb1; t.w[i][30] = mb;
        t.w[i][31] = 0u;
    }
    return t;
}

template <int R> struct RunImage { alignas(16) uint32_t v[RunTab<R>::WORDS]; };
constexpr uint32_t run_tab_permuted(uint32_t e)                    // prmt(e, 0, 0x0423): bytes [e3, e2, 0, e0]
{
    return (e >> 24) | (((e >> 16) & 255u) << 8) | ((e & 255u) << 24);
}
template <int R>
constexpr RunImage<R> make_run_image()
{
    RunImage<R> t{};
    constexpr int HB = RunTab<R>::HB;
    for (int m = 0; m < RunTab<R>::SIZE; ++m) {
        int rs = 0;
        for (uint32_t w = (uint32_t)m; w; w &= w >> 1)
            for (uint32_t x = w; x; x >>= 1) rs += (int)(x & 1u);
        int trail = 0, lead = 0;
        while (trail < HB && ((m >> trail) & 1)) ++trail;
        while (lead < HB && ((m >> (HB - 1 - lead)) & 1)) ++lead;
        t.v[m] = (uint32_t)rs | ((uint32_t)lead << 8) | ((uint32_t)trail << 16) | (1u << 24);
        if (RunTab<R>::COPIES == 2) t.v[RunTab<R>::SIZE + m] = run_tab_permuted(t.v[m]);
    }
    return t;
}

// 1 << a, opaque to the optimiser (which would turn x * (1 << a) back into a shift)
TB_HD uint32_t pow2_opaque(int a)
{
#if defined(__CUDA_ARCH__)
    uint32_t p;
    asm("shl.b32 %0, 1, %1;" : "=r"(p) : "r"(a));
    return p;
#else
    return 1u << a;
#endif
}

// Incremental evaluation of one placement: orientation `u` (width W) anchored at column c.  Only the piece's
// columns and their neighbours are re-evaluated; everything else comes from the env record.  Branch-free apart
// from the two early exits.
template <int C, int R, int W>
TB_HD int eval_neigh(const uint32_t *rec, const uint32_t *runtab, const Neigh<C, R, W> &nb, const OriU &u, int c, Eval &e)
{
    using K = Env<C, R>;
    int h[W + 2];
    uint32_t y[W + 4];
#pragma unroll
    for (int k = 0; k < W + 2; ++k) h[k] = nb.h[k];
#pragma unroll
    for (int k = 0; k < W + 4; ++k) y[k] = nb.y[k];
    int a = 0;
#pragma unroll
    for (int dx = 0; dx < W; ++dx) a = imax(a, h[1 + dx] - u.bot[dx]);          // tetromino.py: anchor_row = max(h - bottom)

    // Shifts by the anchor row as multiplications by 2^a: the kernels are bound by the half-rate ALU pipe (LOP3 / SHF /
    // IADD3: 67 % busy in K1) while the multiply-add pipe idles at 18 %, so `x << a` = x * p2a and the masks
    // mask_lo(a + k) = p2a * 2^k - 1 each become ONE IMAD on the other pipe (profiles/README.md, r2m).
    const uint32_t p2a = pow2_opaque(a);
    // rows among changed_lines that the placement completes: empty cells before == piece cells added (see build_env)
    const uint32_t full = (rec[K::FQ + u.qa1] & (u.ma * p2a)) | (rec[K::FQ + u.qb1] & (u.mb * p2a));
    const int top = a + u.ph;
    e.a = a; e.full = full;
    if (full != 0u) {
        // stack rows are contiguous, so clearing k rows lowers the tallest column by exactly k
        const int hmax = (int)reinterpret_cast<const uint8_t *>(rec)[K::B_HMAX];
        e.terminal = (imax(hmax, top) - popc32(full)) > R;
        return kFastClears;
    }
    if (top > R) { e.terminal = 1; return kFastTerminal; }
    e.terminal = 0;

    int gapsum = 0, gapcnt = 0, hdadd = 0;
    uint32_t gapor = 0;
    uint32_t mt[W];                                                // mask_lo(new height) of the piece columns
#pragma unroll
    for (int dx = 0; dx < W; ++dx) {
        const int len = u.len[dx], lo = a + u.bot[dx], hh = h[1 + dx];
        const int g = lo - hh;                                     // new holes under the piece in this column
        const int g1 = imin(g, 1);                                 // g > 0 (g is never negative: a = max(h - bot))
        y[2 + dx] += u.seg[dx] * p2a;                              // the cells land on empty cells: OR = ADD
        gapsum += g;
        gapcnt += g1;
        gapor |= (p2a * u.p2bot[dx] - 1u) ^ nb.mh[dx];             // rows hh .. lo-1
        hdadd += len * (nb.nr[dx] + g1);
        h[1 + dx] = lo + len;
        mt[dx] = p2a * u.p2top[dx] - 1u;
    }

    // wells over columns c-1 .. c+W, row transitions over columns c .. c+W; the sentinels make the walls come out right
    uint32_t wells = nb.wells0, rt = nb.rt0;
#pragma unroll
    for (int k = 1; k <= W + 2; ++k) wells = run_sum_acc<R>(runtab, y[k - 1] & y[k + 1] & ~y[k], wells);
    // (skipping the lookups of a column when no lane of the warp has a well cell there -- a vote per column -- was
    //  measured 8 % slower in K1 and 7 % in K3: profiles/README.md, r2c)
#pragma unroll
    for (int dx = 0; dx < W; ++dx) {                               // piece columns: height > 0
        const int hj = h[1 + dx], hl = h[dx];
        rt += (uint32_t)(imax(0, hl - hj) + popc32((y[2 + dx] ^ y[1 + dx]) & mt[dx]));
    }
    {                                                              // right neighbour: a column, or the wall
        const int hj = h[W + 1], hl = h[W];
        const uint32_t m = hj > 0 ? ((y[W + 2] ^ y[W + 1]) & nb.mh[W]) : y[W + 1];
        const int v = popc32(m) + (hj > 0 ? imax(0, hl - hj) : 0);
        rt += (uint32_t)((c + W < C) ? v : -u.len[W - 1]);         // wall term R - popc(col[C-1]) loses the new cells
    }
    e.f[0] = unbias(kFloatBias + (uint32_t)popc32(rec[K::T_HM] | gapor));
    e.f[1] = unbias(rec[K::T_CT] + 2u * (uint32_t)gapcnt);
    e.f[2] = unbias(rec[K::T_HOLES] + (uint32_t)gapsum);
    e.f[3] = unbias(2u * (uint32_t)a + u.lh2) * 0.5f;
    e.f[4] = unbias(wells);
    e.f[5] = unbias(rt);
    e.f[6] = 0.0f;
    e.f[7] = unbias(rec[K::T_HD] + (uint32_t)hdadd);
    return kFastDone;
}

template <int C, int R, int W>
TB_HD int eval_placement(const uint32_t *rec, const uint32_t *runtab, uint32_t d, int c, Eval &e)
{
    Neigh<C, R, W> nb;
    load_neigh<C, R, W>(rec, c, nb);
    return eval_neigh<C, R, W>(rec, runtab, nb, decode_ori(d), c, e);
}

// width dispatch (W is warp-uniform in the kernels)
template <int C, int R>
TB_HD int eval_placement_w(const uint32_t *rec, const uint32_t *runtab, uint32_t d, int c, Eval &e)
{
    switch (desc_w(d)) {
    case 1: return eval_placement<C, R, 1>(rec, runtab, d, c, e);
    case 2: return eval_placement<C, R, 2>(rec, runtab, d, c, e);
    case 3: return eval_placement<C, R, 3>(rec, runtab, d, c, e);
    default: return eval_placement<C, R, 4>(rec, runtab, d, c, e);
    }
}

// Tetris.fitness (game.py:109-120): float32 products and sums, left to right, no FMA contraction.
TB_HD float fitness(const float *f, const float *w)
{
#if defined(__CUDA_ARCH__)
    float acc = __fmul_rn(f[0], w[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) acc = __fadd_rn(acc, __fmul_rn(f[i], w[i]));
    return acc;
#else
    volatile float acc = f[0] * w[0];
    for (int i = 1; i < 8; ++i) { volatile float p = f[i] * w[i]; acc = acc + p; }
    return acc;
#endif
}

// Is the placement a legal action, i.e. a non-terminal afterstate (game.py:69)?  Stack rows are contiguous, so
// clearing k rows lowers the tallest column by exactly k: terminal <=> max(hmax, a + piece height) - k > R.
// hmax = tallest column of the current board.  Precondition: hmax <= R.
template <int C, int R>
TB_HD bool placement_valid(const uint32_t *col, uint32_t d, int c, int hmax)
{
    using S = Shape<C, R>;
    const int a = anchor_from_cols<C>(col, d, c);
    const int top = imax(hmax, a + desc_ph(d));
    if (top <= R) return true;
    uint32_t full = S::ALL;
    if (C <= 12) {                                               // branch-free over the columns, as in place_and_clear
        const unsigned long long q = (unsigned long long)piece_cells4(d) << (4 * c);
#pragma unroll
        for (int k = 0; k < C; ++k) full &= col[k] | (((uint32_t)(q >> (4 * k)) & 15u) << a);
    } else {
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const unsigned dx = (unsigned)(k - c);
            uint32_t x = col[k];
            if (dx < 4u) {
                const uint32_t f = (d >> (3 + 5 * dx)) & 31u;
                x |= mask_lo((int)(f >> 2)) << (a + (int)(f & 3u));
            }
            full &= x;
        }
    }
    full &= mask_lo(desc_chg(d)) << a;
    return top - popc32(full) <= R;
}

// Legal-slot mask of a piece on a board (game.py:69 over the enumeration of tetromino.py), one thread per env,
// from column heights alone wherever that decides it: a placement whose top stays at or below row R is legal; one
// that reaches above R is legal only if it clears enough rows, which needs a row that is full but for <= 4 cells
// (`nearfull`, a bit-sliced count of the empty cells of every row) -- only then is the exact test run.
// `ori` = the orientation table (shared memory on the device).  Precondition: every column height <= R.
// ANY_ONLY: return non-zero as soon as one legal slot is found (the mask is then not complete).
// insert a zero above every one of the low 16 bits
TB_HD uint32_t spread16(uint32_t x)
{
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// bit i of the low half -> bit 2i, bit i of the high half -> bit 2i + 1 (outer perfect shuffle): the loop-local slot bits
// c * 2 + o of a column loop with two orientations, from the orientations' column masks packed as two 16-bit halves
TB_HD uint32_t interleave16(uint32_t x)
{
    uint32_t t;
    t = (x ^ (x >> 8)) & 0x0000FF00u; x ^= t ^ (t << 8);
    t = (x ^ (x >> 4)) & 0x00F000F0u; x ^= t ^ (t << 4);
    t = (x ^ (x >> 2)) & 0x0C0C0C0Cu; x ^= t ^ (t << 2);
    t = (x ^ (x >> 1)) & 0x22222222u; x ^= t ^ (t << 1);
    return x;
}

// Per orientation, for the column-parallel legality test of valid_slots: a piece column dx pokes above row R when its
// board column is taller than R - 4 + k, k = 4 - ph + bot[dx] in 0..3.  With the four height classes G_0..G_3 packed as the
// 16-bit fields of a 64-bit word, field k is picked by ONE byte permute; this table holds its selector for every dx
// (16 bits each: bytes 2k, 2k+1 into both halves of the result).
struct PokeImage { unsigned long long v[kNumOris]; };
constexpr PokeImage make_poke_image()
{
    PokeImage t{};
    for (int i = 0; i < kNumOris; ++i) {
        const uint32_t d = kOriHost[i];
        const int ph = (int)((d >> 28) & 7u);
        unsigned long long v = 0;
        for (int dx = 0; dx < 4; ++dx) {
            const int k = (4 - ph + (int)((d >> (3 + 5 * dx)) & 3u)) & 3;   // columns beyond the piece's width: unused
            v |= (unsigned long long)(2u * (uint32_t)k * 0x1111u + 0x1010u) << (16 * dx);
        }
        t.v[i] = v;
    }
    return t;
}
static constexpr PokeImage kPokeHost = make_poke_image();
#if defined(__CUDACC__)
static __constant__ PokeImage c_poke = make_poke_image();
#endif
TB_HD unsigned long long poke_sel(int oi)
{
#if defined(__CUDA_ARCH__)
    return c_poke.v[oi];
#else
    return kPokeHost.v[oi];
#endif
}

template <int C, int R, bool ANY_ONLY = false>
TB_HD unsigned long long valid_slots(const uint32_t *col, uint32_t pw, const uint32_t *ori)
{
    using S = Shape<C, R>;
    const int n0 = pw & 3, w0 = (pw >> 2) & 7, n1 = (pw >> 5) & 3, w1 = (pw >> 7) & 7, obase = (pw >> 10) & 63;
    const int total = n0 * (C - w0 + 1) + n1 * (C - w1 + 1);
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < C; ++k) any |= col[k];
    const int hmax = height_of(any);
    if (hmax + 4 <= R) return (1ull << total) - 1ull;   // every piece is at most 4 rows tall
    if (C > 4) {
        // Tall board, all columns at once.  A placement that stays at or below row R is legal.  One that pokes above R
        // (anchor a >= R - 3) is legal only if it completes enough rows among a .. R-1 (rows >= R are empty).  With
        // G_k = {columns taller than R - 4 + k} as a bit mask, "top <= R" <=> every piece column dx has
        // h[c + dx] <= R - ph + bot[dx] = R - 4 + k (k = 4 - ph + bot is always in 0..3), so the anchor columns that poke
        // above R are OR_dx (G_k(dx) >> dx): no loop over slots.  A poking placement can only be rescued by completing a
        // row r in R-3 .. R-1 that has at most 4 empty cells, all of them under the piece: its anchor column then lies in
        // [hi_r - w + 1, lo_r] (lo_r / hi_r = first / last empty cell of the row) -- only those few get the exact test.
        // The top four legal rows, as row masks, come from the columns by a 4 x C bit gather.
        // (bit j of a column's top nibble -> bit 0 of byte j by one multiply; columns 0-7 and 8-9 in two accumulators)
        uint32_t acc0 = 0u, acc1 = 0u;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const uint32_t x = ((col[k] >> (R - 4)) * 0x00204081u) & 0x01010101u;   // heights are <= R: at most 4 bits
            if (k < 8) acc0 |= x << k; else acc1 |= x << (k - 8);
        }
        uint32_t row[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) row[j] = ((acc0 >> (8 * j)) & 0xFFu) | (((acc1 >> (8 * j)) & 0xFFu) << 8);
        const uint32_t g3 = row[3], g2 = g3 | row[2], g1 = g2 | row[1], g0 = g1 | row[0];
        const uint32_t gg_lo = g0 | g1 << 16, gg_hi = g2 | g3 << 16;  // G_0..G_3 as 16-bit fields (poke_sel picks one)
        // near-full rows among R-3 .. R-1: span of their empty cells, lo | hi << 8 | 1 << 16 (0 = not near-full)
        uint32_t span[3] = {0u, 0u, 0u};
        const bool near = popc32(row[1]) >= C - 4 || popc32(row[2]) >= C - 4 || popc32(row[3]) >= C - 4;
        if (near) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint32_t e = ~row[j + 1] & S::FULLROW;
                if (popc32(e) <= 4 && e != 0u) span[j] = (uint32_t)ctz32(e) | (uint32_t)(31 - clz32(e)) << 8 | 1u << 16;
            }
        }
        unsigned long long m = 0ull, cs = 0ull;                // legal slots; poking slots that need the exact test
#pragma unroll 1
        for (int l = 0; l < 2; ++l) {                          // rolled on purpose: instruction footprint
            const int n = l ? n1 : n0, w = l ? w1 : w0, ob = l ? obase + n0 : obase;
            const int sbase = l ? n0 * (C - w0 + 1) : 0;
            const uint32_t range = mask_lo(C - w + 1);
            uint32_t rescue = 0u;                              // anchor columns whose piece covers a near-full row's gap
            if (near) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int lo = (int)(span[j] & 0xFFu), hi = (int)((span[j] >> 8) & 0xFFu);
                    if (span[j] != 0u && hi - lo < w) rescue |= mask_lo(lo + 1) & ~mask_lo(imax(hi - w + 1, 0));
                }
            }
            uint32_t okp = 0u, cdp = 0u;                       // per orientation: 16 bits of legal / to-be-tested columns
#pragma unroll 1
            for (int o = 0; o < n; ++o) {
                const unsigned long long ps = poke_sel(ob + o);
                uint32_t bad = 0u;
#pragma unroll
                for (int dx = 0; dx < 4; ++dx)
                    if (dx < w) {
                        uint32_t g = prmt(gg_lo, gg_hi, (uint32_t)(ps >> (16 * dx)));   // field k(dx), in both halves
                        if (C > 12) g &= 0xFFFFu;              // narrower boards: `range` cuts what the shift lets in
                        bad |= g >> dx;
                    }
                const uint32_t ok = ~bad & range, cd = bad & rescue & range;
                if (ANY_ONLY && ok != 0u) return 1ull;
                okp |= ok << (16 * o); cdp |= cd << (16 * o);
            }
            // loop-local slot of (column c, orientation o) = c * n + o
            m |= (unsigned long long)(n == 2 ? interleave16(okp) : okp) << sbase;
            if (cdp != 0u) cs |= (unsigned long long)(n == 2 ? interleave16(cdp) : cdp) << sbase;
        }
        if (ANY_ONLY && m != 0ull) return 1ull;
        while (cs != 0ull) {                                   // rare: the one copy of the exact test
            const uint32_t lo32 = (uint32_t)cs;
            const int slot = lo32 ? ctz32(lo32) : 32 + ctz32((uint32_t)(cs >> 32));
            cs &= cs - 1ull;
            int oi, c;
            slot_to_placement(pw, C, slot, oi, c);
            if (placement_valid<C, R>(col, ori[oi], c, hmax)) {
                if (ANY_ONLY) return 1ull;
                m |= 1ull << slot;
            }
        }
        return m;
    }
    uint64_t hp = 0;                                    // heights, 5 bits per column
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;            // bit-sliced count of empty cells per row
#pragma unroll
    for (int k = 0; k < C; ++k) {
        hp |= (uint64_t)height_of(col[k]) << (5 * k);
        const uint32_t e = ~col[k] & S::ALL;
        const uint32_t k0 = c0 & e; c0 ^= e;
        const uint32_t k1 = c1 & k0; c1 ^= k0;
        const uint32_t k2 = c2 & k1; c2 ^= k1;
        c3 |= k2;
    }
    const uint32_t nearfull = ~(c3 | (c2 & (c1 | c0)));  // rows with at most 4 empty cells
    unsigned long long m = 0ull;
    for (int l = 0; l < 2; ++l) {
        const int n = l ? n1 : n0, w = l ? w1 : w0, ob = l ? obase + n0 : obase;
        const int sbase = l ? n0 * (C - w0 + 1) : 0;
        for (int o = 0; o < n; ++o) {
            const uint32_t d = ori[ob + o];
            const int b0 = desc_bot(d, 0), b1 = desc_bot(d, 1), b2 = desc_bot(d, 2), b3 = desc_bot(d, 3);
            const int ph = desc_ph(d);
            const uint32_t chgm = mask_lo(desc_chg(d));
            for (int c = 0; c + w <= C; ++c) {
                const uint32_t hw = (uint32_t)(hp >> (5 * c));
                int a = (int)(hw & 31u) - b0;
                if (w > 1) a = imax(a, (int)((hw >> 5) & 31u) - b1);
                if (w > 2) a = imax(a, (int)((hw >> 10) & 31u) - b2);
                if (w > 3) a = imax(a, (int)((hw >> 15) & 31u) - b3);
                bool ok = imax(hmax, a + ph) <= R;
                if (!ok && ((nearfull >> a) & chgm) != 0u) ok = placement_valid<C, R>(col, d, c, hmax);
                if (ANY_ONLY && ok) return 1ull;        // is_game_over only asks whether a legal placement exists
                m |= (unsigned long long)ok << (sbase + c * n + o);
            }
        }
    }
    return m;
}

// State.__init__ on a caller-supplied board (state.py:5-38): clear the full rows among the `chg` changed lines
// starting at row a, terminal test, features.  ppcr = pieces_per_changed_row packed 4 bits each.
template <int C, int R>
TB_HD void eval_state(uint32_t *col, int a, int chg, uint32_t ppcr, int bonus2, Eval &e)
{
    using S = Shape<C, R>;
    uint32_t full = S::ALL;
#pragma unroll
    for (int k = 0; k < C; ++k) full &= col[k];
    full &= mask_lo(chg) << a;
    int eroded = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < chg && ((full >> (a + k)) & 1u)) eroded += (int)((ppcr >> (4 * k)) & 15u);
    if (full) clear_rows<C>(col, full);
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < C; ++k) any |= col[k];
    int six[6];
    eval_full<C, R>(col, six);
    e.a = a; e.full = full; e.terminal = (int)((any >> R) & 1u);
    e.f[0] = (float)six[0]; e.f[1] = (float)six[1]; e.f[2] = (float)six[2];
    e.f[3] = (float)(2 * (a + 1) + bonus2) * 0.5f;
    e.f[4] = (float)six[3]; e.f[5] = (float)six[4];
    e.f[6] = (float)(eroded * popc32(full));
    e.f[7] = (float)six[5];
}

}  // namespace tb
