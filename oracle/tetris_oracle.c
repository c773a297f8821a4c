/*
 * tetris_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, cell-by-cell CPU restatement of the s0phia-/tetris reference
 * algorithm for the hot path (afterstate enumeration, line clearing, terminal
 * test, the eight BCTS features, the env step loop).  It deliberately follows
 * the reference's *loops over cells* (one byte per cell, one int per column
 * height) and none of the bit-parallel forms the CUDA kernels use, so that
 * agreement between the two is evidence and not a tautology.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (tetris_b200/) never does and fails loudly without its CUDA extension.
 *
 * Parity pin: the reference ships no tests or golden vectors,
 * so this oracle is pinned against fixtures generated from the live Python
 * reference by tests/golden/make_golden.py (committed together with the
 * fixtures) -- see tests/test_oracle_golden.py.
 *
 * Reference citations are file:line into the reference checkout
 * (game.py, state.py, tetromino.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define ORC_MAX_N 32   /* stored rows  = num_rows + 4 (game.py:56)          */
#define ORC_MAX_C 16   /* columns                                             */
#define ORC_MAX_A 64   /* afterstates per piece (36 at C=10, 60 at C=16: ThreeL) */

/* ------------------------------------------------------------------------ */
/* Piece x orientation tables, transcribed from tetromino.py.                */
/* A piece has 1-2 column loops; inside a loop the listed orientations are   */
/* emitted interleaved per column (tetromino.py e.g. :347-378).              */
/* ------------------------------------------------------------------------ */
typedef struct {
    int w;               /* width in columns                                  */
    int ncells;
    int cell[4][2];      /* (dx, dy) relative to (anchor_col, anchor_row)     */
    int chg;             /* len(changed_lines): rows a .. a+chg-1             */
    int ppcr[4];         /* pieces_per_changed_row                            */
    float bonus;         /* landing_height_bonus                              */
} orc_ori;

typedef struct {
    const char *name;
    int nloops;
    int loop_n[2];       /* orientations in loop 0 / loop 1                   */
    orc_ori ori[4];      /* loop 0's orientations first, then loop 1's        */
} orc_piece;

/* Global piece ids.  0..6 = the 7-piece order of game.py:41-47,             */
/* 7,8 = the default two-piece set of game.py:38-39 (ThreeL, ThreeLine).     */
enum { P_STRAIGHT = 0, P_RCORNER, P_LCORNER, P_SQUARE, P_SNAKER, P_SNAKEL, P_T,
       P_THREEL, P_THREELINE, ORC_NPIECES };

static const orc_piece PIECES[ORC_NPIECES] = {
    /* Straight  tetromino.py:33-75 */
    { "Straight", 2, {1, 1}, {
        { 1, 4, {{0,0},{0,1},{0,2},{0,3}}, 4, {1,1,1,1}, 1.5f },   /* :44-57 */
        { 4, 4, {{0,0},{1,0},{2,0},{3,0}}, 1, {4,0,0,0}, 0.0f },   /* :60-74 */
    } },
    /* RCorner   tetromino.py:417-495 */
    { "RCorner", 2, {2, 2}, {
        { 3, 4, {{0,0},{1,0},{2,0},{2,1}}, 1, {3,0,0,0}, 0.5f },   /* :431-445 */
        { 3, 4, {{0,0},{0,1},{1,1},{2,1}}, 2, {1,3,0,0}, 0.5f },   /* :447-460 */
        { 2, 4, {{0,2},{1,0},{1,1},{1,2}}, 3, {1,1,2,0}, 1.0f },   /* :465-478 */
        { 2, 4, {{0,0},{0,1},{0,2},{1,0}}, 1, {2,0,0,0}, 1.0f },   /* :480-494 */
    } },
    /* LCorner   tetromino.py:498-576 */
    { "LCorner", 2, {2, 2}, {
        { 3, 4, {{0,0},{1,0},{2,0},{0,1}}, 1, {3,0,0,0}, 0.5f },   /* :511-525 */
        { 3, 4, {{2,0},{0,1},{1,1},{2,1}}, 2, {1,3,0,0}, 0.5f },   /* :527-540 */
        { 2, 4, {{0,0},{0,1},{0,2},{1,2}}, 3, {1,1,2,0}, 1.0f },   /* :545-559 */
        { 2, 4, {{0,0},{1,0},{1,1},{1,2}}, 1, {2,0,0,0}, 1.0f },   /* :561-575 */
    } },
    /* Square    tetromino.py:78-104 */
    { "Square", 1, {1, 0}, {
        { 2, 4, {{0,0},{1,0},{0,1},{1,1}}, 2, {2,2,0,0}, 0.5f },   /* :90-103 */
    } },
    /* SnakeR    tetromino.py:107-154 */
    { "SnakeR", 2, {1, 1}, {
        { 3, 4, {{0,0},{1,0},{1,1},{2,1}}, 1, {2,0,0,0}, 0.5f },   /* :120-135 */
        { 2, 4, {{0,1},{0,2},{1,0},{1,1}}, 2, {1,2,0,0}, 1.0f },   /* :138-153 */
    } },
    /* SnakeL    tetromino.py:285-331 */
    { "SnakeL", 2, {1, 1}, {
        { 3, 4, {{1,0},{2,0},{0,1},{1,1}}, 1, {2,0,0,0}, 0.5f },   /* :297-312 */
        { 2, 4, {{0,0},{0,1},{1,1},{1,2}}, 2, {1,2,0,0}, 1.0f },   /* :315-330 */
    } },
    /* T         tetromino.py:334-414 */
    { "T", 2, {2, 2}, {
        { 3, 4, {{0,0},{1,0},{2,0},{1,1}}, 1, {3,0,0,0}, 0.5f },   /* :348-362 */
        { 3, 4, {{1,0},{0,1},{1,1},{2,1}}, 2, {1,3,0,0}, 0.5f },   /* :364-378 */
        { 2, 4, {{0,1},{1,0},{1,1},{1,2}}, 2, {1,2,0,0}, 1.0f },   /* :383-397 */
        { 2, 4, {{0,0},{0,1},{0,2},{1,1}}, 2, {1,2,0,0}, 1.0f },   /* :399-413 */
    } },
    /* ThreeL    tetromino.py:202-282 */
    { "ThreeL", 2, {2, 2}, {
        { 2, 3, {{0,0},{1,0},{1,1},{0,0}}, 1, {2,0,0,0}, 0.5f },   /* :216-230 */
        { 2, 3, {{0,0},{0,1},{1,1},{0,0}}, 2, {1,2,0,0}, 0.5f },   /* :232-247 */
        { 2, 3, {{0,1},{1,0},{1,1},{0,0}}, 2, {1,2,0,0}, 0.5f },   /* :252-266 */
        { 2, 3, {{0,0},{0,1},{1,0},{0,0}}, 1, {2,0,0,0}, 0.5f },   /* :267-281 */
    } },
    /* ThreeLine tetromino.py:157-199 */
    { "ThreeLine", 2, {1, 1}, {
        { 1, 3, {{0,0},{0,1},{0,2},{0,0}}, 3, {1,1,1,0}, 1.0f },   /* :168-181 */
        { 3, 3, {{0,0},{1,0},{2,0},{0,0}}, 1, {3,0,0,0}, 0.0f },   /* :184-198 */
    } },
};

/* piece sets: 0 = reference default (game.py:38-39), 1 = 7-piece (game.py:41-47) */
static const int SET_N[2] = { 2, 7 };
static const int SET_PIECES[2][7] = {
    { P_THREEL, P_THREELINE, 0, 0, 0, 0, 0 },
    { P_STRAIGHT, P_RCORNER, P_LCORNER, P_SQUARE, P_SNAKER, P_SNAKEL, P_T },
};

int orc_set_size(int piece_set) { return SET_N[piece_set]; }
int orc_set_piece(int piece_set, int idx) { return SET_PIECES[piece_set][idx]; }
const char *orc_piece_name(int piece) { return PIECES[piece].name; }

int orc_num_afterstates(int piece, int C)
{
    const orc_piece *p = &PIECES[piece];
    int n = 0, o = 0;
    for (int l = 0; l < p->nloops; ++l)
        for (int k = 0; k < p->loop_n[l]; ++k, ++o)
            n += C - p->ori[o].w + 1;
    return n;
}

/* ------------------------------------------------------------------------ */
/* state.py                                                                  */
/* ------------------------------------------------------------------------ */

/* state.py:162-172 calc_lowest_free_rows */
void orc_calc_lowest_free_rows(int C, int N, const uint8_t *rep, int *h)
{
    for (int c = 0; c < C; ++c) {
        int lowest = 0;
        for (int r = N - 1; r >= 0; --r)
            if (rep[r * C + c] == 1) { lowest = r + 1; break; }
        h[c] = lowest;
    }
}

/* state.py:121-143 clear_lines_jitted.  changed = a .. a+chg-1.            */
static int clear_lines(int C, int N, uint8_t *rep, int *h, int a, int chg, uint8_t *is_full)
{
    int lines_to_clear[4], n = 0;
    for (int k = 0; k < chg; ++k) {
        int sum = 0;
        for (int c = 0; c < C; ++c) sum += rep[(a + k) * C + c];        /* :122 */
        is_full[k] = (sum == C);                                        /* :123 */
        if (is_full[k]) lines_to_clear[n++] = a + k;
    }
    if (n > 0) {
        uint8_t keep[ORC_MAX_N];
        uint8_t tmp[ORC_MAX_N * ORC_MAX_C];
        for (int r = 0; r < N; ++r) keep[r] = 1;
        for (int k = 0; k < n; ++k) keep[lines_to_clear[k]] = 0;        /* :128-129 */
        int w = 0;
        for (int r = 0; r < N; ++r)
            if (keep[r]) { memcpy(tmp + w * C, rep + r * C, (size_t)C); ++w; }
        for (; w < N; ++w) memset(tmp + w * C, 0, (size_t)C);           /* :130-131 */
        memcpy(rep, tmp, (size_t)(N * C));
        for (int c = 0; c < C; ++c) {                                   /* :132-142 */
            int old = h[c];
            if (old > lines_to_clear[n - 1] + 1) {
                h[c] -= n;
            } else {
                int lowest = 0;
                for (int r = old - n - 1; r >= 0; --r)
                    if (rep[r * C + c] == 1) { lowest = r + 1; break; }
                h[c] = lowest;
            }
        }
    }
    return n;
}

/* state.py:111-117 check_terminal: any cell in row index n_legal_rows.      */
static int check_terminal(int C, int R, const uint8_t *rep)
{
    for (int c = 0; c < C; ++c)
        if (rep[R * C + c]) return 1;
    return 0;
}

/* state.py:175-280 get_feature_values_jitted -- cell loop, walls of ones,   */
/* wall heights = num_rows (n_legal_rows).  out = [rows_with_holes,          */
/* column_transitions, holes, cumulative_wells, row_transitions, hole_depth] */
static void feature_values(int C, int R, int N, const uint8_t *rep, const int *h, int *out)
{
    /* rep_x[r][0] and rep_x[r][C+1] are the walls (state.py:177-178) */
    uint8_t x[ORC_MAX_N][ORC_MAX_C + 2];
    int hx[ORC_MAX_C + 2];
    for (int r = 0; r < N; ++r) {
        x[r][0] = 1; x[r][C + 1] = 1;
        for (int c = 0; c < C; ++c) x[r][c + 1] = rep[r * C + c];
    }
    hx[0] = R; hx[C + 1] = R;                                           /* :179 */
    for (int c = 0; c < C; ++c) hx[c + 1] = h[c];

    uint8_t row_has_hole[ORC_MAX_N];
    memset(row_has_hole, 0, sizeof row_has_hole);
    int column_transitions = 0, holes = 0, cumulative_wells = 0;
    int row_transitions = 0, hole_depth = 0;

    { int s = 0; for (int r = 0; r < N; ++r) s += x[r][C]; row_transitions += R - s; } /* :190 */

    for (int ci = 1; ci <= C; ++ci) {                                   /* :192 */
        int lfr = hx[ci];
        column_transitions += 1;                                        /* :194 */
        int streak = 0;
        if (lfr > 0) {                                                  /* :197 */
            int full_above = 0;
            for (int r = 0; r < lfr; ++r) full_above += x[r][ci];       /* :200 */
            if (hx[ci - 1] > hx[ci]) row_transitions += hx[ci - 1] - hx[ci]; /* :203-204 */
            int cell_below = 1;
            for (int r = 0; r < lfr; ++r) {                             /* :208 */
                int cell = x[r][ci];
                if (cell == 0) {
                    holes += 1;                                         /* :213 */
                    row_has_hole[r] = 1;                                /* :215 */
                    if (r + 1 < lfr && x[r + 1][ci] == 1) hole_depth += full_above; /* :216 */
                    if (cell_below) column_transitions += 1;            /* :219-220 */
                    int left = x[r][ci - 1], right = x[r][ci + 1];
                    if (left) {
                        row_transitions += 1;                           /* :225-226 */
                        if (right) { streak += 1; cumulative_wells += streak; } /* :227-229 */
                        else streak = 0;
                    } else streak = 0;
                } else {
                    streak = 0;                                         /* :236 */
                    full_above -= 1;                                    /* :239 */
                    if (!cell_below) column_transitions += 1;           /* :242-243 */
                    if (!x[r][ci - 1]) row_transitions += 1;            /* :246-248 */
                }
                cell_below = cell;
            }
        } else {
            int s = 0;
            for (int r = 0; r < hx[ci - 1]; ++r) s += x[r][ci - 1];
            row_transitions += s;                                       /* :254 */
        }
        int lim = hx[ci - 1] < hx[ci + 1] ? hx[ci - 1] : hx[ci + 1];    /* :258-260 */
        if (lim > lfr) {
            for (int r = lfr; r < lim; ++r) {                           /* :262-272 */
                if (x[r][ci - 1]) {
                    if (x[r][ci + 1]) { streak += 1; cumulative_wells += streak; }
                    else streak = 0;
                } else streak = 0;
            }
        }
    }
    int rwh = 0;
    for (int r = 0; r < N; ++r) rwh += row_has_hole[r];                 /* :274-275 */
    out[0] = rwh; out[1] = column_transitions; out[2] = holes;
    out[3] = cumulative_wells; out[4] = row_transitions; out[5] = hole_depth;
}

/* An afterstate as State.__init__ leaves it (state.py:5-38). */
typedef struct {
    uint8_t rep[ORC_MAX_N * ORC_MAX_C];
    int h[ORC_MAX_C];
    int anchor_col, anchor_row, n_cleared, terminal, chg;
    uint8_t is_full[4];          /* cleared_rows_relative_to_anchor            */
    int ppcr[4];
    float bonus;
    float feat[8];               /* calc_bcts_features state.py:97-107         */
} orc_state;

static void calc_features(int C, int R, int N, orc_state *s)
{
    int eroded = 0, ncl = 0, six[6];
    for (int k = 0; k < s->chg; ++k) { eroded += s->is_full[k] * s->ppcr[k]; ncl += s->is_full[k]; } /* :99-100 */
    feature_values(C, R, N, s->rep, s->h, six);
    s->feat[6] = (float)(eroded * ncl);                                 /* :101 */
    s->feat[3] = (float)s->anchor_row + s->bonus + 1.0f;                /* :102 */
    s->feat[0] = (float)six[0]; s->feat[1] = (float)six[1]; s->feat[2] = (float)six[2];
    s->feat[4] = (float)six[3]; s->feat[5] = (float)six[4]; s->feat[7] = (float)six[5]; /* :103 */
}

/* The empty state Tetris.reset() builds (game.py:55-58): default changed_lines=[0],
 * pieces_per_changed_row=[0], bonus 0 (state.py:7-9). */
static void make_reset_state(int C, int R, int N, orc_state *s)
{
    memset(s, 0, sizeof *s);
    s->chg = 1;
    s->n_cleared = clear_lines(C, N, s->rep, s->h, 0, 1, s->is_full);
    s->terminal = check_terminal(C, R, s->rep);
    calc_features(C, R, N, s);
}

void orc_reset_state_features(int C, int R, float *feat /*[8]*/)
{
    orc_state s;
    make_reset_state(C, R, R + 4, &s);
    memcpy(feat, s.feat, sizeof s.feat);
}

/* tetromino.py <Piece>.get_after_states: enumerate in reference order. */
static int enumerate(int C, int R, int piece, const uint8_t *rep, const int *h, orc_state *out, int with_features)
{
    const int N = R + 4;
    const orc_piece *p = &PIECES[piece];
    int n = 0, obase = 0;
    for (int l = 0; l < p->nloops; ++l) {
        /* all orientations of one loop share the same width (tetromino.py max_col_index) */
        int w = p->ori[obase].w;
        for (int c = 0; c + w <= C; ++c) {
            for (int k = 0; k < p->loop_n[l]; ++k) {
                const orc_ori *o = &p->ori[obase + k];
                int bot[4] = {99, 99, 99, 99}, top[4] = {-1, -1, -1, -1};
                for (int i = 0; i < o->ncells; ++i) {
                    int dx = o->cell[i][0], dy = o->cell[i][1];
                    if (dy < bot[dx]) bot[dx] = dy;
                    if (dy > top[dx]) top[dx] = dy;
                }
                int a = -99;
                for (int dx = 0; dx < o->w; ++dx)
                    if (h[c + dx] - bot[dx] > a) a = h[c + dx] - bot[dx];
                orc_state *s = &out[n++];
                memcpy(s->rep, rep, (size_t)(N * C));
                memcpy(s->h, h, sizeof(int) * (size_t)C);
                for (int i = 0; i < o->ncells; ++i)
                    s->rep[(a + o->cell[i][1]) * C + c + o->cell[i][0]] = 1;
                for (int dx = 0; dx < o->w; ++dx) s->h[c + dx] = a + top[dx] + 1;
                s->anchor_col = c;
                s->anchor_row = a;                                      /* state.py:32 */
                s->chg = o->chg;
                memcpy(s->ppcr, o->ppcr, sizeof s->ppcr);
                s->bonus = o->bonus;
                memset(s->is_full, 0, sizeof s->is_full);
                s->n_cleared = clear_lines(C, N, s->rep, s->h, a, o->chg, s->is_full); /* state.py:33 */
                s->terminal = check_terminal(C, R, s->rep);             /* state.py:36 */
                if (with_features) calc_features(C, R, N, s);
            }
        }
        obase += p->loop_n[l];
    }
    return n;
}

/* ------------------------------------------------------------------------ */
/* Exported single-board entry points (used by the parity tests).            */
/* ------------------------------------------------------------------------ */

/* Enumerate all afterstates of `piece` on one board.  Output arrays are sized
 * for ORC_MAX_A afterstates; returns the number written. */
int orc_afterstates(int C, int R, int piece, const uint8_t *rep, const int32_t *h_in,
                    float *feats /*[A][8]*/, uint8_t *terminal /*[A]*/, int32_t *n_cleared /*[A]*/,
                    uint8_t *rep_out /*[A][N*C]*/, int32_t *h_out /*[A][C]*/,
                    int32_t *anchor /*[A][2] col,row*/, uint8_t *is_full /*[A][4]*/)
{
    const int N = R + 4;
    orc_state st[ORC_MAX_A];
    int h[ORC_MAX_C];
    for (int c = 0; c < C; ++c) h[c] = h_in[c];
    int n = enumerate(C, R, piece, rep, h, st, 1);
    for (int i = 0; i < n; ++i) {
        if (feats) memcpy(feats + 8 * i, st[i].feat, sizeof st[i].feat);
        if (terminal) terminal[i] = (uint8_t)st[i].terminal;
        if (n_cleared) n_cleared[i] = st[i].n_cleared;
        if (rep_out) memcpy(rep_out + (size_t)i * N * C, st[i].rep, (size_t)(N * C));
        if (h_out) for (int c = 0; c < C; ++c) h_out[i * C + c] = st[i].h[c];
        if (anchor) { anchor[2 * i] = st[i].anchor_col; anchor[2 * i + 1] = st[i].anchor_row; }
        if (is_full) memcpy(is_full + 4 * i, st[i].is_full, 4);
    }
    return n;
}

/* Features of an arbitrary board as State(representation) would report them
 * (state.py:5-38 with default changed_lines=[0], ppcr=[0], bonus=0). */
void orc_board_features(int C, int R, const uint8_t *rep_in, float *feat /*[8]*/,
                        int32_t *h_out, uint8_t *rep_out, int32_t *terminal, int32_t *n_cleared)
{
    const int N = R + 4;
    orc_state s;
    memset(&s, 0, sizeof s);
    memcpy(s.rep, rep_in, (size_t)(N * C));
    orc_calc_lowest_free_rows(C, N, s.rep, s.h);
    s.chg = 1;
    s.n_cleared = clear_lines(C, N, s.rep, s.h, 0, 1, s.is_full);
    s.terminal = check_terminal(C, R, s.rep);
    calc_features(C, R, N, &s);
    memcpy(feat, s.feat, sizeof s.feat);
    if (h_out) for (int c = 0; c < C; ++c) h_out[c] = s.h[c];
    if (rep_out) memcpy(rep_out, s.rep, (size_t)(N * C));
    if (terminal) *terminal = s.terminal;
    if (n_cleared) *n_cleared = s.n_cleared;
}

/* Tetris.fitness game.py:109-120: float32 products and sums, left to right
 * (NumPy >= 2: np.float32 * python float -> np.float32). */
static float fitness_w(const float *f, const float *w)
{
    volatile float acc = f[0] * w[0];
    for (int i = 1; i < 8; ++i) { volatile float p = f[i] * w[i]; acc = acc + p; }
    return acc;
}
float orc_fitness(const float *f, const float *w) { return fitness_w(f, w); }

/* ------------------------------------------------------------------------ */
/* Per-env counter-based RNG + shuffled bag.  This is NOT reference code: the */
/* reference draws from NumPy's global MT19937 (tetromino.py:15,19), which a  */
/* per-env in-kernel generator cannot reproduce; this is the CPU restatement  */
/* of the generator the CUDA kernels use (tetris_b200/csrc/tb_core.cuh), used */
/* to drive oracle traces and to inject identical piece tapes into the        */
/* reference through its replaceable `tetromino_sampler` attribute.           */
/* Bag semantics follow TetrominoSampler (tetromino.py:12-22): a permutation  */
/* is consumed front to back and refilled when empty.                         */
/* ------------------------------------------------------------------------ */
static uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint32_t orc_rng(uint64_t seed, uint64_t env, uint32_t ctr, uint32_t stream)
{
    uint64_t k = mix64(seed + 0x9E3779B97F4A7C15ULL * (env + 1));
    uint64_t z = mix64(k ^ (((uint64_t)stream << 32) | (uint64_t)ctr));
    return (uint32_t)(z >> 32);
}
static uint32_t bounded(uint32_t r, uint32_t k) { return (uint32_t)(((uint64_t)r * (uint64_t)k) >> 32); }

/* Draw the next local piece index from the env's bag. */
static int bag_draw(int n_set, uint64_t seed, uint64_t env, uint32_t *bag, uint32_t *draws)
{
    if (*bag == 0) *bag = (1u << n_set) - 1u;
    uint32_t k = (uint32_t)__builtin_popcount(*bag);
    uint32_t j = bounded(orc_rng(seed, env, *draws, 0u), k);
    *draws += 1;
    uint32_t b = *bag;
    for (uint32_t i = 0; i < j; ++i) b &= b - 1;
    int idx = __builtin_ctz(b);
    *bag &= ~(1u << idx);
    return idx;
}

/* ------------------------------------------------------------------------ */
/* Batched env state (plain arrays, one slot per env) and the env loop.       */
/* ------------------------------------------------------------------------ */
typedef struct {
    int32_t C, R, piece_set;
    int64_t n_env, env_offset;
    uint64_t seed;
    uint8_t *rep;        /* [n][N*C] */
    int32_t *h;          /* [n][C]   */
    int32_t *piece;      /* [n] global piece id */
    uint32_t *bag;       /* [n] */
    uint32_t *draws;     /* [n] */
    uint32_t *ep_steps;  /* [n] */
    uint32_t *ep_lines;  /* [n] */
} orc_batch;

orc_batch *orc_batch_new(int C, int R, int piece_set, int64_t n_env, int64_t env_offset, uint64_t seed)
{
    orc_batch *b = (orc_batch *)calloc(1, sizeof *b);
    const int N = R + 4;
    b->C = C; b->R = R; b->piece_set = piece_set; b->n_env = n_env; b->env_offset = env_offset; b->seed = seed;
    b->rep = (uint8_t *)calloc((size_t)n_env * N * C, 1);
    b->h = (int32_t *)calloc((size_t)n_env * C, sizeof(int32_t));
    b->piece = (int32_t *)calloc((size_t)n_env, sizeof(int32_t));
    b->bag = (uint32_t *)calloc((size_t)n_env, sizeof(uint32_t));
    b->draws = (uint32_t *)calloc((size_t)n_env, sizeof(uint32_t));
    b->ep_steps = (uint32_t *)calloc((size_t)n_env, sizeof(uint32_t));
    b->ep_lines = (uint32_t *)calloc((size_t)n_env, sizeof(uint32_t));
    return b;
}
void orc_batch_free(orc_batch *b)
{
    if (!b) return;
    free(b->rep); free(b->h); free(b->piece); free(b->bag); free(b->draws); free(b->ep_steps); free(b->ep_lines); free(b);
}
uint8_t *orc_batch_rep(orc_batch *b) { return b->rep; }
int32_t *orc_batch_heights(orc_batch *b) { return b->h; }
int32_t *orc_batch_piece(orc_batch *b) { return b->piece; }
uint32_t *orc_batch_bag(orc_batch *b) { return b->bag; }
uint32_t *orc_batch_draws(orc_batch *b) { return b->draws; }

static void env_reset_board(orc_batch *b, int64_t e)
{
    const int C = b->C, N = b->R + 4;
    memset(b->rep + (size_t)e * N * C, 0, (size_t)(N * C));
    memset(b->h + (size_t)e * C, 0, sizeof(int32_t) * (size_t)C);
}
static void env_draw(orc_batch *b, int64_t e, const uint8_t *tape)
{
    if (tape) { b->piece[e] = tape[e]; b->draws[e] += 1; return; }
    int idx = bag_draw(SET_N[b->piece_set], b->seed, (uint64_t)(b->env_offset + e), &b->bag[e], &b->draws[e]);
    b->piece[e] = SET_PIECES[b->piece_set][idx];
}

/* Tetris.__init__ + reset (game.py:21-63): empty board, fresh bag, draw one piece.
 * tape (nullable) = global piece id per env to use instead of the bag RNG. */
void orc_batch_reset(orc_batch *b, const uint8_t *tape)
{
    for (int64_t e = 0; e < b->n_env; ++e) {
        env_reset_board(b, e);
        b->bag[e] = 0; b->draws[e] = 0; b->ep_steps[e] = 0; b->ep_lines[e] = 0;
        env_draw(b, e, tape);
    }
}

/* Tetris.get_after_states(include_terminal=True) game.py:67-80 for every env:
 * features of every afterstate by enumeration slot, valid (= non-terminal)
 * mask, count of valid.  a_max = row stride of feats. */
void orc_batch_afterstates(orc_batch *b, int a_max, float *feats /*[n][a_max][8]*/,
                           uint64_t *valid /*[n]*/, int32_t *count /*[n]*/, int32_t *n_all /*[n]*/)
{
    const int C = b->C, R = b->R, N = R + 4;
    for (int64_t e = 0; e < b->n_env; ++e) {
        orc_state st[ORC_MAX_A];
        int h[ORC_MAX_C];
        for (int c = 0; c < C; ++c) h[c] = b->h[e * C + c];
        int n = enumerate(C, R, b->piece[e], b->rep + (size_t)e * N * C, h, st, 1);
        uint64_t m = 0; int cnt = 0;
        for (int i = 0; i < n; ++i) {
            if (feats) memcpy(feats + ((size_t)e * a_max + i) * 8, st[i].feat, sizeof st[i].feat);
            if (!st[i].terminal) { m |= 1ull << i; ++cnt; }
        }
        if (valid) valid[e] = m;
        if (count) count[e] = cnt;
        if (n_all) n_all[e] = n;
    }
}

/* One env: Tetris.step(action) game.py:82-92 (+ optional auto-reset as the
 * example_play.py:20-21 caller would do).  action = index into the
 * non-terminal afterstates (game.py:69,83), or an enumeration slot if
 * action_is_slot.  Returns 0, or -1 for an out-of-range action (IndexError). */
static int env_step(orc_batch *b, int64_t e, int action, int action_is_slot, const uint8_t *tape,
                    int auto_reset, float *obs, int32_t *reward, uint8_t *done, int32_t *lines)
{
    const int C = b->C, R = b->R, N = R + 4;
    orc_state st[ORC_MAX_A];
    int h[ORC_MAX_C];
    uint8_t *rep = b->rep + (size_t)e * N * C;
    for (int c = 0; c < C; ++c) h[c] = b->h[e * C + c];
    int n = enumerate(C, R, b->piece[e], rep, h, st, 0);
    int sel = -1;
    if (action_is_slot) {
        if (action >= 0 && action < n && !st[action].terminal) sel = action;
    } else {
        int k = 0;
        for (int i = 0; i < n; ++i)
            if (!st[i].terminal) { if (k == action) { sel = i; break; } ++k; }
    }
    if (sel < 0 || action < 0) return -1;
    orc_state *s = &st[sel];
    memcpy(rep, s->rep, (size_t)(N * C));                                /* game.py:83 */
    for (int c = 0; c < C; ++c) b->h[e * C + c] = s->h[c];
    int lc = s->n_cleared;                                              /* :85 */
    int rew = lc - 1;                                                   /* :86 timestep_reward */
    env_draw(b, e, tape);                                               /* :87 */
    /* is_game_over game.py:94-100: second enumeration, new piece on new board */
    orc_state st2[ORC_MAX_A];
    int n2 = enumerate(C, R, b->piece[e], rep, s->h, st2, 0);
    int any = 0;
    for (int i = 0; i < n2; ++i) if (!st2[i].terminal) { any = 1; break; }
    int dn = !any;
    if (dn) rew += -100;                                                /* :89-90 loss_reward */
    if (obs) { calc_features(C, R, N, s); memcpy(obs, s->feat, sizeof s->feat); } /* :91 */
    if (reward) *reward = rew;
    if (done) *done = (uint8_t)dn;
    if (lines) *lines = lc;
    b->ep_steps[e] += 1; b->ep_lines[e] += (uint32_t)lc;
    if (dn && auto_reset) {                                             /* example_play.py:20-21 -> game.py:53-63 */
        env_reset_board(b, e);
        env_draw(b, e, NULL);   /* with a tape the caller resets explicitly */
        b->ep_steps[e] = 0; b->ep_lines[e] = 0;
    }
    return 0;
}

int orc_batch_step(orc_batch *b, const int32_t *actions, int action_is_slot, const uint8_t *tape, int auto_reset,
                   float *obs /*[n][8]*/, int32_t *reward, uint8_t *done, int32_t *lines)
{
    int bad = 0;
    for (int64_t e = 0; e < b->n_env; ++e) {
        int rc = env_step(b, e, actions[e], action_is_slot, tape, auto_reset && !tape,
                          obs ? obs + 8 * e : NULL, reward ? reward + e : NULL,
                          done ? done + e : NULL, lines ? lines + e : NULL);
        if (rc) bad = 1;
    }
    return bad ? -1 : 0;
}

/* Reset only the envs flagged in `mask` (what example_play.py does on done):
 * board emptied, one more piece drawn (from the tape if given). */
void orc_batch_reset_masked(orc_batch *b, const uint8_t *mask, const uint8_t *tape)
{
    for (int64_t e = 0; e < b->n_env; ++e)
        if (mask[e]) { env_reset_board(b, e); env_draw(b, e, tape); b->ep_steps[e] = 0; b->ep_lines[e] = 0; }
}

/* stats layout shared with the CUDA rollout kernel (include/tetris_b200.h) */
enum { ST_PLACEMENTS = 0, ST_EPISODES, ST_LINES, ST_REWARD, ST_AFTERSTATES,
       ST_LINES0, ST_LINES1, ST_LINES2, ST_LINES3, ST_LINES4,
       ST_MAX_EP_LINES, ST_MAX_EP_STEPS, ST_SUM_EP_STEPS, ST_SUM_EP_LINES, ST_RESERVED0, ST_RESERVED1, ST_COUNT };

/* The example_play.py:11-21 loop with an in-loop policy, T placements per env,
 * auto-reset on done.  policy 0 = uniformly random valid action (policy RNG
 * stream 1, counter = draws); policy 1 = greedy linear: first arg-max over the
 * non-terminal afterstates of the float32 left-to-right score (game.py:109-120). */
static void rollout_range(orc_batch *b, int64_t lo, int64_t hi, int T, int policy, const float *weights, int64_t *loc)
{
    const int C = b->C, R = b->R, N = R + 4;
    for (int64_t e = lo; e < hi; ++e) {
        for (int t = 0; t < T; ++t) {
            orc_state st[ORC_MAX_A];
            int h[ORC_MAX_C];
            for (int c = 0; c < C; ++c) h[c] = b->h[e * C + c];
            int n = enumerate(C, R, b->piece[e], b->rep + (size_t)e * N * C, h, st, policy == 1);
            int nv = 0;
            for (int i = 0; i < n; ++i) nv += !st[i].terminal;
            int action = 0;
            if (policy == 0) {
                action = (int)bounded(orc_rng(b->seed, (uint64_t)(b->env_offset + e), b->draws[e], 1u), (uint32_t)nv);
            } else {
                float best = 0.f; int k = 0, have = 0;
                for (int i = 0; i < n; ++i) {
                    if (st[i].terminal) continue;
                    float f = fitness_w(st[i].feat, weights);
                    if (!have || f > best) { best = f; action = k; have = 1; }
                    ++k;
                }
            }
            loc[ST_AFTERSTATES] += n;
            int32_t rew = 0, lc = 0; uint8_t dn = 0;
            env_step(b, e, action, 0, NULL, 0, NULL, &rew, &dn, &lc);
            uint32_t eps = b->ep_steps[e], epl = b->ep_lines[e];
            loc[ST_PLACEMENTS] += 1; loc[ST_LINES] += lc; loc[ST_REWARD] += rew; loc[ST_LINES0 + lc] += 1;
            if (dn) {
                loc[ST_EPISODES] += 1; loc[ST_SUM_EP_STEPS] += eps; loc[ST_SUM_EP_LINES] += epl;
                if ((int64_t)epl > loc[ST_MAX_EP_LINES]) loc[ST_MAX_EP_LINES] = epl;
                if ((int64_t)eps > loc[ST_MAX_EP_STEPS]) loc[ST_MAX_EP_STEPS] = eps;
                env_reset_board(b, e);
                env_draw(b, e, NULL);
                b->ep_steps[e] = 0; b->ep_lines[e] = 0;
            }
        }
    }
}

typedef struct { orc_batch *b; int64_t lo, hi; int T, policy; const float *weights; int64_t loc[ST_COUNT]; } rollout_job;
static void *rollout_thread(void *arg)
{
    rollout_job *j = (rollout_job *)arg;
    rollout_range(j->b, j->lo, j->hi, j->T, j->policy, j->weights, j->loc);
    return NULL;
}

void orc_batch_rollout_mt(orc_batch *b, int T, int policy, const float *weights, int64_t *stats /*[ST_COUNT]*/, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > b->n_env) nthreads = (int)b->n_env;
    if (nthreads < 1) return;
    rollout_job *jobs = (rollout_job *)calloc((size_t)nthreads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
    for (int i = 0; i < nthreads; ++i) {
        jobs[i].b = b; jobs[i].T = T; jobs[i].policy = policy; jobs[i].weights = weights;
        jobs[i].lo = b->n_env * i / nthreads; jobs[i].hi = b->n_env * (i + 1) / nthreads;
        if (nthreads > 1) pthread_create(&th[i], NULL, rollout_thread, &jobs[i]);
        else rollout_thread(&jobs[i]);
    }
    for (int i = 0; i < nthreads; ++i) {
        if (nthreads > 1) pthread_join(th[i], NULL);
        for (int k = 0; k < ST_COUNT; ++k) {
            if (k == ST_MAX_EP_LINES || k == ST_MAX_EP_STEPS) { if (jobs[i].loc[k] > stats[k]) stats[k] = jobs[i].loc[k]; }
            else stats[k] += jobs[i].loc[k];
        }
    }
    free(jobs); free(th);
}

/* Tetris.perform_rollouts / single_rollout (game.py:129-160) for every env and every action, with the fork RNG
 * convention of tb_rollout_values: child d = (e * a_stride + slot) * n_forks + f is a one-env copy of the parent
 * whose pieces come from the stream (seed2, child_offset + d), continuing the parent's bag.  ret_sum[e][slot] = sum
 * over forks of the rollout return: -1 if the game ended (after the action or on the way), else the sum of the
 * follow-up rewards.  valid[e] = legal slots.
 * piece_tape (nullable): uint8[n_env * a_stride * n_forks][length], the pieces child d draws, in order, instead of
 * its RNG stream -- a recorded sampler of the reference (tests/golden/make_golden.py gen_rollouts). */
void orc_batch_rollout_values(orc_batch *b, int a_stride, int n_forks, int length, int policy, const float *weights,
                              uint64_t seed2, int64_t child_offset, const uint8_t *piece_tape, int32_t *ret_sum,
                              uint64_t *valid)
{
    const int C = b->C, R = b->R, N = R + 4;
    orc_batch *t = orc_batch_new(C, R, b->piece_set, 1, 0, seed2);
    for (int64_t e = 0; e < b->n_env; ++e) {
        orc_state st[ORC_MAX_A];
        int h[ORC_MAX_C];
        for (int c = 0; c < C; ++c) h[c] = b->h[e * C + c];
        const int n = enumerate(C, R, b->piece[e], b->rep + (size_t)e * N * C, h, st, 0);
        uint64_t vm = 0;
        for (int s = 0; s < a_stride; ++s) {
            int sum = 0;
            const int legal = s < n && !st[s].terminal;
            if (legal) vm |= 1ull << s;
            for (int f = 0; legal && f < n_forks; ++f) {
                t->env_offset = child_offset + ((int64_t)e * a_stride + s) * n_forks + f;
                memcpy(t->rep, b->rep + (size_t)e * N * C, (size_t)(N * C));
                memcpy(t->h, b->h + (size_t)e * C, sizeof(int32_t) * (size_t)C);
                t->piece[0] = b->piece[e]; t->bag[0] = b->bag[e]; t->draws[0] = b->draws[e];
                t->ep_steps[0] = 0; t->ep_lines[0] = 0;
                uint8_t dn = 0; int32_t rew = 0, lc = 0;
                const uint8_t *tp = piece_tape ? piece_tape + (((size_t)e * a_stride + s) * n_forks + f) * (size_t)length : NULL;
                env_step(t, 0, s, 1, tp, 0, NULL, &rew, &dn, &lc);                 /* game.py:132 */
                int ret = 0;
                if (dn) ret = -1;                                                   /* :133-137 */
                for (int k = 0; !dn && k < length - 1; ++k) {                       /* :139-145 */
                    orc_state s2[ORC_MAX_A];
                    int h2[ORC_MAX_C];
                    for (int c = 0; c < C; ++c) h2[c] = t->h[c];
                    const int n2 = enumerate(C, R, t->piece[0], t->rep, h2, s2, policy == 1);
                    int nv = 0, action = 0;
                    for (int i = 0; i < n2; ++i) nv += !s2[i].terminal;
                    if (policy == 0) {
                        action = (int)bounded(orc_rng(t->seed, (uint64_t)t->env_offset, t->draws[0], 1u), (uint32_t)nv);
                    } else {
                        float best = 0.f; int kk = 0, have = 0;
                        for (int i = 0; i < n2; ++i) {
                            if (s2[i].terminal) continue;
                            const float fv = fitness_w(s2[i].feat, weights);
                            if (!have || fv > best) { best = fv; action = kk; have = 1; }
                            ++kk;
                        }
                    }
                    env_step(t, 0, action, 0, tp ? tp + 1 + k : NULL, 0, NULL, &rew, &dn, &lc);
                    ret += rew;
                    if (dn) ret = -1;
                }
                sum += ret;
            }
            ret_sum[e * a_stride + s] = sum;
        }
        if (valid) valid[e] = vm;
    }
    orc_batch_free(t);
}

void orc_batch_rollout(orc_batch *b, int T, int policy, const float *weights, int64_t *stats /*[ST_COUNT]*/)
{
    orc_batch_rollout_mt(b, T, policy, weights, stats, 1);
}

uint32_t *orc_batch_ep_steps(orc_batch *b) { return b->ep_steps; }
uint32_t *orc_batch_ep_lines(orc_batch *b) { return b->ep_lines; }
int orc_stats_count(void) { return ST_COUNT; }
